import sys, os, time, gc
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from vamp_b200 import batch, engine, synth
ctx = engine.default_context(0)
reg = synth.config_c2(region_id=2)
regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])])
models = [(0, n) for n in range(1, 7)]
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
def step(seed, keep):
    return batch.fit_models(regs, models, voigt=True, iterations=3000, burn=300, thin=15, restarts=4, chains_per_restart=1024,
                            polish_top=4, dtype="f32", seed=seed, region_uids=[2], keep_sampler=keep)
for keep in (True, False, True, False):
    last = None
    for w in range(2):
        last = step(1000 + w, keep)
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
    t0 = time.perf_counter()
    evs[0].record()
    for k in range(8):
        flush.fill_(float(k))
        last = step(2000 + k, keep)
        evs[k + 1].record()
    torch.cuda.synchronize()
    print("keep_sampler", keep, ["%.0f" % evs[k].elapsed_time(evs[k + 1]) for k in range(8)], "wall %.0f" % ((time.perf_counter() - t0) * 1e3 / 8), flush=True)
