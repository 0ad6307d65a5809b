"""How even is the static LPT partition?  Fits every rank's shard of an N-rank run of R C4 regions, one after the other on
ONE GPU (``fit_regions(emulate=(rank, N))``) and prints the shard times:  python tools/shard_balance.py [R] [N]
Measured (16 384 regions, 8 shards): mean 3.39 s, max 3.46 s - the partition by pixels x free parameters is even to 2 %;
a cost model with a fixed part per Metropolis step was tried and is no better (max / mean 1.03 at 128 pixel-equivalents,
1.10 at 256).  One eighth of the one-GPU time would be 2.97 s: what a small shard loses is the latency-bound end of every
ladder round, not balance."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vamp_b200 import synth
from vamp_b200.fit_regions import fit_regions

R = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
import multiprocessing as mp
with mp.get_context("fork").Pool(16) as pool:
    regs = pool.map(synth.config_c4_region, range(10_000, 10_000 + R), chunksize=256)
host = [(r["nu"], r["flux"], r["noise"]) for r in regs]
uids = [10_000 + i for i in range(R)]
kw = dict(voigt=True, iterations=3000, burn=300, thin=15, restarts=3, chains_per_restart=8, window=(2, 1), nmax=8, seed=1)
fit_regions(host[:512], uids=uids[:512], **kw)           # warm-up
times = []
for rank in range(N):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = fit_regions(host, uids=uids, emulate=(rank, N), polish_dtype="mixed", **kw)
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
    print("rank %d: %d regions %.3f s" % (rank, len(res), times[-1]), flush=True)
t = np.array(times)
print("shard times mean %.3f s, max %.3f s, max / mean %.3f, spread %.1f %%" % (
    t.mean(), t.max(), t.max() / t.mean(), 100 * (t.max() - t.min()) / t.mean()))
