"""One large K1 launch for ncu:  python tools/profile_k1.py [f32|f64]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vamp_b200 import engine, synth, _lib
from vamp_b200._lib import check
dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
ctx = engine.default_context(0); dev = ctx.torch_device
P, C, B = 256, 8, 262144 if dtype == "f32" else 65536
reg = synth.make_region(P, 8, 50.0, 7000 + P + C, voigt=True)
regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])])
rng = np.random.default_rng(5)
th = np.stack([synth.prior_draw_theta(reg["nu"], C, True, rng, posterior_like=reg["truth"]) for _ in range(4096)])
flat = engine.pack_theta(regs.to_scaled(th, np.zeros(4096, np.int32)), None, C)
d_th = torch.from_numpy(np.tile(flat, (B // 4096, 1))).to(dev)
d_ir = torch.zeros(B, dtype=torch.int32, device=dev); d_in = torch.full((B,), C, dtype=torch.int32, device=dev)
d_sd = torch.full((B,), 0.02, dtype=torch.float64, device=dev)
d_lnL = torch.empty(B, dtype=torch.float64, device=dev); d_chi = torch.empty(B, dtype=torch.float64, device=dev)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(ctx.lib.vamp_loglike_dev(ctx.handle, regs.handle, engine.dtype_code(dtype), _lib.VAMP_VOIGT, B, d_ir.data_ptr(), d_in.data_ptr(), C,
                                   d_th.data_ptr(), 4 * C + 1, d_sd.data_ptr(), d_lnL.data_ptr(), d_chi.data_ptr(), None, P))
    e1.record(); torch.cuda.synchronize()
print("K1 %s P %d C %d items %d: %.3f ms  %.3e evals/s" % (dtype, P, C, B, e0.elapsed_time(e1), P * C * B / e0.elapsed_time(e1) * 1e3))
