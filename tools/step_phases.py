"""Phase timing of bench steps with CUDA events (K2 / K3 / rest):  python tools/step_phases.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vamp_b200 import batch, engine, synth

ctx = engine.default_context(0)
reg = synth.config_c2(region_id=2)
regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])])
models = [(0, n) for n in range(1, 7)]
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
orig_run, orig_map = engine.Sampler.run, engine.Sampler._map
marks = []
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
def run(self, iters, *a_, **k_):
    marks.append(("k2_start", ev())); orig_run(self, iters, *a_, **k_); marks.append(("k2_end", ev()))
def _map(self, *a, **k):
    marks.append(("k3_start", ev())); r = orig_map(self, *a, **k); marks.append(("k3_end", ev())); return r
engine.Sampler.run, engine.Sampler._map = run, _map
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    marks.clear()
    flush.fill_(float(k))
    torch.cuda.synchronize()
    t0 = time.perf_counter(); marks.append(("start", ev()))
    mb = batch.fit_models(regs, models, voigt=True, iterations=3000, burn=300, thin=15, restarts=4, chains_per_restart=1024,
                          polish_top=4, dtype="f32", seed=2000 + k, region_uids=[2], keep_sampler=True,
                          polish_dtype=os.environ.get("POLISH", "f64"))
    marks.append(("end", ev())); torch.cuda.synchronize(); wall = time.perf_counter() - t0
    d = dict(marks)
    print("step %d: wall %.1f ms | pre-K2 %.1f  K2 %.1f  K2->K3 %.1f  K3 %.1f  post %.1f | nm iters max %d mean %.0f" % (
        k, wall * 1e3, d["start"].elapsed_time(d["k2_start"]), d["k2_start"].elapsed_time(d["k2_end"]),
        d["k2_end"].elapsed_time(d["k3_start"]), d["k3_start"].elapsed_time(d["k3_end"]), d["k3_end"].elapsed_time(d["end"]),
        mb.nm_iters.max(), mb.nm_iters.mean()))
