"""Sampler rate on long regions:  python tools/k2_time_long.py [iters]   (float32, Voigt, 3 components, 2048 chains per length)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vamp_b200 import engine, synth
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
ctx = engine.default_context(0)
for P in (256, 512, 1024, 2048):
    reg = synth.make_region(P, 3, 50.0, 4000 + P, voigt=True)
    regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])])
    B = 8192
    s = engine.Sampler(regs, np.zeros(B, np.int32), np.full(B, 3, np.int32), np.arange(B, dtype=np.uint64) + 1, voigt=True,
                       dtype="f32", seed=1, burn=30, thin=15, tune_interval=50, tune_throughout=True)
    best = 1e9
    for rep in range(3):
        s.init(); s.run(5); torch.cuda.synchronize()
        w0 = s.work()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.run(iters); e1.record(); torch.cuda.synchronize()
        w1 = s.work()
        best = min(best, e0.elapsed_time(e1))
    ev = w1["shape_pix"] - w0["shape_pix"]
    print("P %4d: %d chains x %d iterations %.1f ms  %.3g evals/s" % (P, B, iters, best, ev / best * 1e3), flush=True)
    regs.close()
