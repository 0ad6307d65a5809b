"""Hand-counted algorithmic work of the shipped kernels (FMA = 2 flops, MUFU = 1).

Used by ``bench.py`` for ``roofline.achieved`` and stated in DESIGN.md section 5; cross-check with
ncu ``smsp__sass_thread_inst_executed_op_{fadd,fmul,ffma}_pred_on`` (profiles/).
Counts follow ``vamp_b200/csrc/vamp_device.cuh`` line by line.

fp32 ``fad_core`` (Weideman N = 12, real second-order recurrence), per pixel:
    r = rcp(s + 2L*y + L*L)            FMA + ADD + MUFU           = 4
    Zr = (L*L - s) * r                 ADD + MUL                  = 2
    Zi = 2L * x * r                    2 MUL                      = 2
    t = 2 Zr, n = |Z|^2                ADD, MUL + FMA             = 4
    10 recurrence steps (2 FMA)        10 * 4                     = 40
    pr (2 FMA), pi (MUL)                                          = 5
    ir, ii                             ADD + 2 MUL                = 3
    i2r, i2i                           MUL + FMA, 2 MUL           = 5
    q, H                               MUL + FMA, FMA + MUL       = 6
                                                                   ---- 71
fp32 ``fad_far`` (4-level continued fraction):
    rs, tr, ti                         MUFU + 2 MUL               = 3
    vr, vi                             MUL + FMA, 2 MUL           = 5
    N(v), D(v): 2 Horner steps each    4 * 7                      = 28
    er, ei, imq                        3 * (MUL + FMA)            = 9
    |D|^2, rcp, -y*imq*rcp             MUL + FMA + MUFU + 2 MUL   = 6
                                                                   ---- 51
common per shape pixel: d = du - cb, x = d*g, s = fma(x,x,y2)     = 4
per likelihood pixel:  v = fma(An,H,base) (tau pre-scaled by -log2 e), ex2 (MUFU), r = f - m, S = fma(r,r,S)  = 6
    (+ 1 for the noise weight, + 2 per LSF tap)
fp64: same structure with N = 32 (30 recurrence steps) and 9 CF levels (4 + 5 Horner steps); the
    reciprocal and exp are multi-instruction sequences, counted as 10 and 30 flops.
"""

F32 = {"core": 71.0, "far": 51.0, "shape_common": 4.0, "like": 6.0}
F64 = {"core": 13 + 2 + 2 + 4 + 30 * 4 + 5 + 3 + 5 + 6, "far": 3 + 5 + 9 * 7 + 9 + 6 + 18, "shape_common": 4.0,
       "like": 36.0}
GAUSS = {"shape": 5.0}      # d, q = d*g, q*q, exp2 (MUFU): SUB + MUL + MUL + MUFU + neg


def k2_flops(work, dtype="f32", voigt=True, weighted=False, lsf_taps=0):
    """Algorithmic flops of the K2 work counters ``{shape_pix, far_pix, like_pix}``."""
    t = F32 if dtype in ("f32", 0) else F64
    like = t["like"] + (1.0 if weighted else 0.0) + (2.0 * lsf_taps if lsf_taps > 1 else 0.0)
    if voigt:
        core_pix = work["shape_pix"] - work["far_pix"]
        shape = core_pix * t["core"] + work["far_pix"] * t["far"] + work["shape_pix"] * t["shape_common"]
    else:
        shape = work["shape_pix"] * GAUSS["shape"]
    return shape + work["like_pix"] * like


def fp32_peak_tflops(sms=148, mhz=1965.0):
    """SMs x 128 lanes x 2 (FMA) x clock."""
    return sms * 128 * 2 * mhz * 1e6 / 1e12


def fp64_peak_tflops(sms=148, mhz=1965.0):
    return sms * 64 * 2 * mhz * 1e6 / 1e12
