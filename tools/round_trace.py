"""Per-round timings of two 2048-region shards of the C4 bench workload (what an 8-GPU run of 16 384 regions gives a rank):
    python tools/round_trace.py
Measured: round 0 (4096 models) K2 2.03 s + polish 0.5-0.6 s; 381 regions go on, the later rounds take 0.1 s each."""
import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from vamp_b200 import synth
from vamp_b200.fit_regions import fit_regions
import multiprocessing as mp
R=16384
with mp.get_context("fork").Pool(16) as pool:
    regs = pool.map(synth.config_c4_region, range(10_000, 10_000 + R), chunksize=256)
host = [(r["nu"], r["flux"], r["noise"]) for r in regs]
uids = [10_000 + i for i in range(R)]
kw = dict(voigt=True, iterations=3000, burn=300, thin=15, restarts=3, chains_per_restart=8, window=(2, 1), nmax=8, seed=1)
fit_regions(host[:512], uids=uids[:512], **kw)
for rank in (0, 1):
    st = {"trace_rounds": True}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st["t0"] = t0
    res = fit_regions(host, uids=uids, emulate=(rank, 8), polish_dtype="mixed", stats=st, **kw)
    torch.cuda.synchronize(); print("rank", rank, "total %.3f s" % (time.perf_counter() - t0), st["timing"])
    prev = (0, 0)
    for r in st["rounds"]:
        print("  round %d: active %5d models %5d nmax %2d  K2 %6d ms  polish %5d ms  (t = %.3f s)" % (r[0], r[1], r[2], r[3], r[4] - prev[0], r[5] - prev[1], r[6]))
        prev = (r[4], r[5])
