"""Hand-counted algorithmic work of the shipped kernels (FMA = 2 flops, MUFU = 1).

Used by ``bench.py`` for ``roofline.achieved`` and stated in DESIGN.md section 5; cross-check with
ncu ``smsp__sass_thread_inst_executed_op_{fadd,fmul,ffma}_pred_on`` (profiles/).
Counts follow ``vamp_b200/csrc/vamp_device.cuh`` line by line.

fp32 ``fad_core2`` (Weideman N = 12, real second-order recurrence; K2's packed two-row form), per pixel:
    R = rcp(s + (2Ly + L^2))           ADD + MUFU                 = 2     (constants: once per component)
    t = K2 R - 2, Zr = K2/2 R - 1, -|Z|^2 = K3 R - 1   3 FMA      = 6
    XR = x R, ir = (L + y) R           2 MUL                      = 2
    b1 = t a0 + a1                     FMA                        = 2
    9 recurrence steps (2 FMA)         9 * 4                      = 36
    pr                                 2 FMA                      = 4
    XR^2, ir^2 - XR^2                  MUL, FMA                   = 3
    (b1 ir) XR^2                       2 MUL                      = 2
    q = -4L v + pr i2r                 MUL + FMA                  = 3
    H = y (sqrt(pi) 2 q + ir)          FMA + MUL                  = 3
                                                                   ---- 63
    (round 1's form, VAMP_FAD_CORE35=0: 64 flops in 40 packed operations; this one: 63 in 35, 27 of them FMAs)
fp32 ``fad_far2`` (4-level continued fraction as N(v)/D(v), v = 1/z^2; both quadratics with real coefficients):
    rs, tr, ti                         MUFU + 2 MUL               = 3
    vr, vi                             MUL + FMA, ADD + MUL       = 5
    a = vr^2 - vi^2                    MUL + FMA                  = 3
    Re N, Im N, Re D, Im D             2 * (2 FMA + FMA + MUL)    = 14
    er, ei, imq                        3 * (MUL + FMA)            = 9
    |D|^2, rcp, -y*imq*rcp             MUL + FMA + MUFU + 2 MUL   = 6
                                                                   ---- 40
    (round 1's form, VAMP_FAD_FAR26=0: two complex Horner steps, 43 flops in 28 packed operations; this one: 40 in 26)
common per shape pixel: d = du - cb, x = d*g, s = fma(x,x,y2)     = 4
per likelihood pixel:  v = fma(An,H,base) (tau pre-scaled by -log2 e), ex2 (MUFU), r = f - m, S = fma(r,r,S)  = 6
    (+ 1 for the noise weight, + 2 per LSF tap)
fp64: same structure with N = 32 (29 recurrence steps) and 9 CF levels (N: 1 real + 3 complex steps, D: 1 real +
    4 complex steps); a reciprocal and an exp are multi-instruction sequences, counted as 10 and 30 flops.
(K1 uses the same packed forms + one FMA per component for the tau accumulation; K3 evaluates one row at a
 time with the scalar forms ``fad_core`` / ``fad_far``: + 7 / + 8.)
"""

F32 = {"core": 63.0, "far": 40.0, "shape_common": 4.0, "like": 6.0}
F64 = {"core": (2 + 9) + 6 + 2 + 2 + 29 * 4 + 4 + 3 + 2 + 3 + 3, "far": (3 + 9) + 5 + (3 + 3 * 7) + (3 + 4 * 7) + 9 + (6 + 9),
       "shape_common": 4.0, "like": 36.0}
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
