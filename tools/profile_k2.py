"""Short K2 (and K3) run on the bench workload shape, for ncu:  python tools/profile_k2.py [iters]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vamp_b200 import batch, engine, synth  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dtype = sys.argv[2] if len(sys.argv) > 2 else "f32"
import torch  # noqa: E402

ctx = engine.default_context(0)
reg = synth.config_c2(region_id=2)
regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])])
mb = batch.fit_models(regs, [(0, n) for n in range(1, 7)], voigt=True, iterations=iters, burn=iters // 10, thin=15,
                      restarts=4, chains_per_restart=1024, dtype=dtype, seed=1, region_uids=[2], keep_sampler=True)
torch.cuda.synchronize()
w = mb.work
print("iters", iters, "shape_pix %.4g far %.3f like_pix %.4g" % (w["shape_pix"], w["far_pix"] / w["shape_pix"], w["like_pix"]),
      "bic", np.round(mb.bic.mean(axis=1), 1))
