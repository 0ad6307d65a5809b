"""Where a bench step goes: python tools/step_breakdown.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vamp_b200 import batch, engine, synth

ctx = engine.default_context(0)
reg = synth.config_c2(region_id=2)
regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])])
models = [(0, n) for n in range(1, 7)]

def step(seed):
    return batch.fit_models(regs, models, voigt=True, iterations=3000, burn=300, thin=15, restarts=4,
                            chains_per_restart=1024, polish_top=4, dtype="f32", seed=seed, region_uids=[2], keep_sampler=True)
for w in range(2):
    mb = step(w)
torch.cuda.synchronize()
import cProfile, pstats
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
mb = step(5)
torch.cuda.synchronize()
pr.disable()
print("wall step %.1f ms" % (1e3 * (time.perf_counter() - t0)))
print("nm iters", mb.nm_iters.tolist())
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
# kernel-level: K2 vs K3
s = mb.sampler
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
ev[0].record(); s.init(); ev[1].record(); s.run(3000); ev[2].record()
torch.cuda.synchronize()
print("init %.2f ms, K2 %.1f ms" % (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
S = 96
starts = s.t["best_theta"][:S].clone()
ev[3].record()
res = s._map(starts, s.t["item_region"][:S].contiguous(), s.t["item_ncomp"][-S:].contiguous(), 6, 3000, 0, 1e-4, 1e-3, "f64")
ev[4].record(); torch.cuda.synchronize()
print("K3 (96 starts, n=6, f64) %.1f ms; iters" % ev[3].elapsed_time(ev[4]), res["iters"].cpu().numpy()[:8])
ev[3].record()
res = s._map(starts, s.t["item_region"][:S].contiguous(), s.t["item_ncomp"][-S:].contiguous(), 6, 3000, 0, 1e-4, 1e-3, "f32")
ev[4].record(); torch.cuda.synchronize()
print("K3 (96 starts, n=6, f32) %.1f ms; iters" % ev[3].elapsed_time(ev[4]), res["iters"].cpu().numpy()[:8])
