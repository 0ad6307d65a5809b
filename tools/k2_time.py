"""Time the K2 launch of the bench workload (C2) with CUDA events:  python tools/k2_time.py [iters] [doublet]
(``doublet``: the same chain set on a CIV doublet region - evals are then LINE evaluations, two per component)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vamp_b200 import batch, engine, synth
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
doublet = len(sys.argv) > 2 and sys.argv[2] == "doublet"       # a CIV doublet region: two tied lines per component
ctx = engine.default_context(0)
reg = synth.make_doublet_region(256, 3, 50.0, 2, "CIV", True) if doublet else synth.config_c2(region_id=2)
regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])], multiplet=reg["multiplet"] if doublet else None)
mb = batch.fit_models(regs, [(0, n) for n in range(1, 7)], voigt=True, iterations=50, burn=5, thin=15, restarts=4,
                      chains_per_restart=1024, dtype="f32", seed=1, region_uids=[2], keep_sampler=True)
s = mb.sampler
best = 1e9
for rep in range(3):
    s.init(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = s.work()
    e0.record(); s.run(iters); e1.record(); torch.cuda.synchronize()
    w1 = s.work()
    ms = e0.elapsed_time(e1); best = min(best, ms)
ev = w1["shape_pix"] - w0["shape_pix"]
print("K2 %d iters: %.2f ms  %.4g evals/s  far %.3f  mean best_logp %.3f" % (iters, best, ev / best * 1e3, (w1["far_pix"] - w0["far_pix"]) / max(ev, 1.0), float(s.t["best_logp"].mean())))
