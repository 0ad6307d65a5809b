import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from vamp_b200 import batch, engine, synth
ctx = engine.default_context(0)
reg = synth.make_region(256, 8, 80.0, 321, voigt=True)
regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])])
for dtype in ("f32", "f64"):
    for n in (8, 12, 15):
        mb = batch.fit_models(regs, [(0, n)], voigt=True, iterations=400 if dtype == "f32" else 60, burn=50, thin=5, restarts=2,
                              chains_per_restart=8, polish_top=2, dtype=dtype, seed=3, maxiter=200, polish_dtype="f64")
        print(dtype, "n", n, "bic", np.round(mb.bic, 1), "chi2/dof", np.round(mb.chi2 / 230, 2), flush=True)
# long run, default mcmc_fit length
mb = batch.fit_models(regs, [(0, 8)], voigt=True, iterations=15000, burn=100, thin=15, restarts=3, chains_per_restart=32,
                      polish_top=4, dtype="f32", seed=5)
print("15000 its n=8: bic", np.round(mb.bic, 1), "chi2/dof", np.round(mb.chi2 / 230, 2), "std finite", bool(np.isfinite(mb.std[0, :, :8]).all()))
