import os, sys, cProfile, pstats
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from vamp_b200 import engine, synth
from vamp_b200.fit_regions import fit_regions
R = 256
regs = [synth.config_c4_region(10_000 + i) for i in range(R)]
host = [(r["nu"], r["flux"], r["noise"]) for r in regs]
kw = dict(uids=[10_000 + i for i in range(R)], voigt=True, iterations=3000, burn=300, thin=15, restarts=3, chains_per_restart=8, window=3, nmax=8)
fit_regions(host, seed=0, **kw); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
fit_regions(host, seed=1, **kw); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
