"""K1 (fused likelihood) throughput sweep - BASELINE config C5 (likelihood only):
pixels x components x parameter vectors, fp32 vs fp64, posterior-like vs prior-draw parameters.
    python tools/bench_k1.py [--quick] [--json out.json]
Prints one line per case: evals/s, algorithmic TFLOP/s (vamp_b200/flops.py counts) and HBM GB/s."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vamp_b200 import engine, flops, synth, _lib
from vamp_b200._lib import check

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--full", action="store_true", help="the whole C5 grid: P x C x chains = 5 x 6 x 4 shapes, both precisions")
ap.add_argument("--json", default=None)
args = ap.parse_args()
ctx = engine.default_context(0)
dev = ctx.torch_device
rng = np.random.default_rng(5)


def case(P, C, B, dtype, mode):
    n_true = max(1, min(C, P // 32, 8))
    reg = synth.make_region(P, n_true, 50.0, 7000 + P + C, voigt=True)
    regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])])
    nb = min(B, 4096)                                   # distinct parameter vectors, tiled to B items
    th = np.stack([synth.prior_draw_theta(reg["nu"], C, True, rng, posterior_like=reg["truth"] if mode == "post" else None)
                   for _ in range(nb)])
    ths = regs.to_scaled(th, np.zeros(nb, np.int32))
    flat = engine.pack_theta(ths, None, C)
    d_th = torch.from_numpy(np.tile(flat, ((B + nb - 1) // nb, 1))[:B].copy()).to(dev)
    d_ir = torch.zeros(B, dtype=torch.int32, device=dev)
    d_in = torch.full((B,), C, dtype=torch.int32, device=dev)
    d_sd = torch.full((B,), 0.02, dtype=torch.float64, device=dev)
    d_lnL = torch.empty(B, dtype=torch.float64, device=dev)
    d_chi = torch.empty(B, dtype=torch.float64, device=dev)
    code = engine.dtype_code(dtype)

    def run():
        check(ctx.lib.vamp_loglike_dev(ctx.handle, regs.handle, code, _lib.VAMP_VOIGT, B, d_ir.data_ptr(), d_in.data_ptr(), C,
                                       d_th.data_ptr(), 4 * C + 1, d_sd.data_ptr(), d_lnL.data_ptr(), d_chi.data_ptr(), None, P))
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    evals = float(P) * C * B
    t = flops.F32 if dtype == "f32" else flops.F64
    # path mix of the kernel's pair policy (all 64 pixels of a row pair at |z|^2 >= R2 -> continued fraction), evaluated
    # on the host for a sample of the parameter vectors
    npair = (P + 63) // 64
    ns = max(1, min(nb, 64, (1 << 23) // (C * npair * 64)))
    u = (reg["nu"] - regs.nu0[0]) / regs.scale[0]
    up = np.full(npair * 64, np.nan)
    up[:P] = u
    up = up.reshape(npair, 64)
    cen, g = ths[:ns, :, 1], 1.66510922231539551270 / ths[:ns, :, 3]
    y2 = (0.83255461115769775635 * ths[:ns, :, 2] / ths[:ns, :, 3]) ** 2
    dmin = np.nanmin(np.abs(up[None, None] - cen[:, :, None, None]), axis=3)                   # [ns, C, npair]
    far_frac = float(np.mean((dmin * g[:, :, None]) ** 2 + y2[:, :, None] >= (32.0 if dtype == "f32" else 64.0)))
    f_mix = evals * (far_frac * t["far"] + (1 - far_frac) * t["core"] + t["shape_common"] + 2) + float(P) * B * (t["like"] + 1)
    f_lo = evals * (t["far"] + t["shape_common"] + 2) + float(P) * B * (t["like"] + 1)
    f_hi = evals * (t["core"] + t["shape_common"] + 2) + float(P) * B * (t["like"] + 1)
    byts = B * (P * (12 if dtype == "f32" else 24) + 32 * C + 16)
    r = {"P": P, "C": C, "items": B, "dtype": dtype, "params": mode, "ms": best, "evals_per_s": evals / best * 1e3,
         "tflops_lo": f_lo / best / 1e9, "tflops_hi": f_hi / best / 1e9, "tflops": f_mix / best / 1e9, "far_path_fraction": far_frac,
         "frac_of_peak": f_mix / best / 1e9 / (flops.fp32_peak_tflops() if dtype == "f32" else flops.fp64_peak_tflops()),
         "hbm_gbs": byts / best / 1e6,
         "lnL_finite": float(torch.isfinite(d_lnL).float().mean())}
    print("P %6d C %2d items %6d %s %-5s: %8.3f ms  %.3e evals/s  %.1f TFLOP/s (%.2f of peak, far %.2f)  %.1f GB/s  finite %.2f" % (
        P, C, B, dtype, mode, best, r["evals_per_s"], r["tflops"], r["frac_of_peak"], far_frac, r["hbm_gbs"], r["lnL_finite"]), flush=True)
    regs.close()
    return r


grid = [(256, 1, 65536), (256, 8, 65536), (256, 32, 16384), (4096, 8, 4096), (65536, 8, 256), (65536, 32, 64), (256, 1, 1), (65536, 1, 1)]
modes = ("post", "prior")
if args.quick:
    grid = [(256, 8, 65536), (4096, 8, 4096)]
if args.full:
    # BASELINE config C5 / SURVEY section 8d: pixels 256-65536 x components 1-32 x chains 1-65536, fp32 vs fp64, the
    # launch capped at P x C x chains <= 2^37 (fp64: 2^35) pixel-component evaluations: 120 shapes per precision
    grid = [(P, C, B) for P in (256, 1024, 4096, 16384, 65536) for C in (1, 2, 4, 8, 16, 32) for B in (1, 64, 4096, 65536)]
    modes = ("post",)
out = []
for P, C, B in grid:
    for dtype in ("f32", "f64"):
        for mode in modes:
            cap = (2 ** 35 if dtype == "f64" else 2 ** 37) if args.full else (2 ** 30 * (4 if dtype == "f32" else 1))
            Bq = B
            while Bq > 1 and Bq * P * C > cap:
                Bq //= 2
            r = case(P, C, Bq, dtype, mode)
            r["items_requested"] = B
            out.append(r)
if args.json:
    json.dump(out, open(args.json, "w"), indent=1)
