"""Small K1/K2/K3/K4 runs for compute-sanitizer:  compute-sanitizer --tool memcheck python tools/sanitize_k2.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vamp_b200 import batch, engine, synth
from vamp_b200 import vpspectrum as VS
ctx = engine.default_context(0)
sizes = [1, 2, 33, 64, 65, 129, 200, 256]
regl = [synth.make_region(P, max(1, min(3, P // 24)), 40.0, 70 + P, voigt=True) for P in sizes]
for lsf in (None, synth.gaussian_lsf(2.0, 7)):
    regs = engine.Regions(ctx, [(r["nu"], r["flux"], r["noise"]) for r in regl], lsf=lsf)
    for sd_free in (True, False):
        for dtype in ("f32", "f64"):
            models = [(i, n) for i in range(len(sizes)) for n in (1, 2, 5) if n <= max(1, (sizes[i] - 1) // 2) or n == 1]
            mb = batch.fit_models(regs, models, voigt=True, sd_free=sd_free, iterations=40, burn=10, thin=3, restarts=2,
                                  chains_per_restart=3, polish_top=2, dtype=dtype, seed=1, maxiter=60,
                                  polish_dtype="mixed" if (dtype == "f32" and lsf is None) else "f64")
            print("lsf", lsf is not None, "sd_free", sd_free, dtype, "bic finite:", float(np.isfinite(mb.bic).mean()), flush=True)
    th = np.tile(regl[-1]["truth"][None], (5, 1, 1))
    for dtype in ("f32", "f64"):
        lnL, chi = engine.loglike(regs, th, np.full(5, len(sizes) - 1, np.int32), sd=np.full(5, 0.03), voigt=True, dtype=dtype)
        print("k1", dtype, lnL[:2])
det = VS.detection_ratio_batch([np.linspace(1000, 1010, P) for P in (5, 300, 700)], [np.full(P, 0.7) for P in (5, 300, 700)],
                               [np.full(P, 0.05) for P in (5, 300, 700)], ctx=ctx)
print("k4", [float(np.nanmax(d[1:-1])) for d in det])
torch.cuda.synchronize()
print("done")
