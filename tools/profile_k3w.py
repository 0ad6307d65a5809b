"""Round 0 of a 2048-region C4 shard (4096 models): the K3w launch groups of its polish, for ncu:
    ncu --set full --clock-control none --import-source on -k regex:k3w -c 4 -o gpurun_out/k3w_r2 python tools/profile_k3w.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vamp_b200 import synth
from vamp_b200.fit_regions import fit_regions
R = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
regs = [synth.config_c4_region(10_000 + i) for i in range(R)]
host = [(r["nu"], r["flux"], r["noise"]) for r in regs]
st = {}
fit_regions(host, uids=[10_000 + i for i in range(R)], voigt=True, iterations=3000, burn=300, thin=15, restarts=3,
            chains_per_restart=8, window=(2, 1), nmax=8, seed=1, max_rounds=1, polish_dtype="mixed", stats=st)
torch.cuda.synchronize()
print(st.get("timing"))
