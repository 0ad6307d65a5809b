import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
from vamp_b200 import engine, synth
ctx = engine.default_context(0)
P, n = 200, 2
reg = synth.make_region(P, n, 50.0, 4321, voigt=True)
regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])])
t = reg["truth"]
print("truth (A, c_px, L_px, G_px):", t[:, 0], (t[:, 1] - reg["nu"][0]) / regs.scale[0], t[:, 2] / regs.scale[0], t[:, 3] / regs.scale[0])
N = 4000
for which, span in ((1, 0.5), (3, 0.5), (2, 0.5), (0, 0.02)):
    th = np.tile(t[None], (N, 1, 1)).copy()
    d = np.linspace(-span, span, N)
    if which == 0:
        th[:, 0, 0] *= (1 + d)
    else:
        th[:, 0, which] += d * regs.scale[0]
    out = {}
    for dt in ("f32", "f64"):
        lnL, chi = engine.loglike(regs, th, np.zeros(N, np.int32), sd=np.full(N, 0.02), voigt=True, dtype=dt)
        out[dt] = lnL
    diff = out["f32"] - out["f64"]
    jumps = np.abs(np.diff(diff))
    print("param %d: lnL range %.1f | f32-f64: mean %.2e std %.2e max|.| %.2e | max jump between neighbours %.2e (99%%: %.2e)" % (
        which, np.ptp(out["f64"]), diff.mean(), diff.std(), np.abs(diff).max(), jumps.max(), np.percentile(jumps, 99)))
