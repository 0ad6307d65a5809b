"""Randomised step-for-step check of the fp64 sampler against the CPU replay (oracle/chain_replay.py):
    python tools/fuzz_replay.py [cases] [seed]
Every case draws a region length (17 .. 2048 pixels: one-tile, pair-tail and long-region kernels), a component count,
Gauss / Voigt, sd free / per-pixel noise, LSF on / off, burn / thin / tune settings and a launch split, runs a few
chains in float64 and compares traces, final states and log-likelihoods with the replay.  Several regions of different
lengths share a region set in every third case (mixed launches).  Exit code 1 on the first mismatch."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import chain_replay as CR
from vamp_b200 import engine, synth

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2024)
ctx = engine.default_context(0)
worst = 0.0
t_start = time.perf_counter()
for case in range(cases):
    voigt = bool(rng.integers(0, 2))
    sd_free = bool(rng.integers(0, 2))
    use_lsf = rng.random() < 0.25
    lsf = synth.gaussian_lsf(float(rng.uniform(1.5, 3.0)), int(rng.choice([3, 7, 11]))) if use_lsf else None
    nreg = 3 if case % 3 == 2 else 1
    Ps = [int(rng.choice([rng.integers(17, 65), rng.integers(65, 257), rng.integers(257, 513), rng.integers(513, 2049)],
                         p=[0.25, 0.35, 0.2, 0.2])) for _ in range(nreg)]
    ns = [int(rng.integers(1, 3 if P < 64 else 5)) for P in Ps]
    regs_h = [synth.make_region(P, n, float(rng.choice([20.0, 50.0])), 5000 + 10 * case + i, voigt=voigt, lsf=lsf)
              for i, (P, n) in enumerate(zip(Ps, ns))]
    regs = engine.Regions(ctx, [(r["nu"], r["flux"], r["noise"]) for r in regs_h], lsf=lsf)
    B = 2 * nreg
    item_region = np.repeat(np.arange(nreg, dtype=np.int32), 2)
    item_ncomp = np.array([ns[r] for r in item_region], dtype=np.int32)
    uid = rng.integers(1, 2 ** 40, B).astype(np.uint64)
    seed = int(rng.integers(1, 2 ** 48))
    iters = int(rng.integers(20, 70))
    burn, thin, tune = int(rng.integers(0, iters // 2)), int(rng.integers(1, 6)), int(rng.integers(5, 30))
    cmax = max(ns) + int(rng.integers(0, 2))
    # rows sorted by (long class, ncomp) as batch.fit_models does, launched per long class
    lc = engine.long_class(regs.npix[item_region])
    order = np.lexsort((np.arange(B), item_ncomp, lc))
    s = engine.Sampler(regs, item_region[order], item_ncomp[order], uid[order], voigt=voigt, sd_free=sd_free, dtype="f64",
                       seed=seed, burn=burn, thin=thin, tune_interval=tune, tune_throughout=True,
                       trace_cap=max(1, (iters - burn) // thin + 2), cmax=cmax)
    s.init()
    th0 = s.get("theta").copy()
    cut = int(rng.integers(1, iters))
    for a, b in ((0, cut), (cut, iters)):
        slc = lc[order]
        for c in np.unique(slc):
            lo, hi = int(np.searchsorted(slc, c, "left")), int(np.searchsorted(slc, c, "right"))
            s.run(b - a, rows=(lo, hi), max_ncomp=int(item_ncomp[order][lo:hi].max()), advance=False)
        s.iteration += b - a
    th, tr, ntal, lnL = s.get("theta"), s.get("trace"), s.get("ntally"), s.get("lnL")
    for row in range(B):
        r, n = int(item_region[order][row]), int(item_ncomp[order][row])
        u = (np.asarray(regs_h[r]["nu"]) - regs.nu0[r]) / regs.scale[r]
        ref = CR.replay_chain(u, regs_h[r]["flux"], regs_h[r]["noise"], th0[row], n, voigt, sd_free, int(uid[order][row]), seed,
                              iters, burn, thin, tune, True, cmax, lsf=lsf)
        ok = ntal[row] == ref["ntally"]
        if ok and ntal[row]:
            d = np.abs(tr[row, : ntal[row]] - ref["trace"]) / (1e-9 + 1e-8 * np.abs(ref["trace"]))
            worst = max(worst, float(d.max()))
            ok = ok and d.max() <= 1.0
        d = np.abs(th[row] - ref["theta"]) / (1e-9 + 1e-8 * np.abs(ref["theta"]))
        ok = ok and d.max() <= 1.0 and abs(lnL[row] - ref["lnL"]) <= 1e-6 + 1e-9 * abs(ref["lnL"])
        if not ok:
            print("MISMATCH case %d row %d: P %s n %s voigt %s sd_free %s lsf %s iters %d burn %d thin %d tune %d cut %d" % (
                case, row, Ps, ns, voigt, sd_free, use_lsf, iters, burn, thin, tune, cut))
            sys.exit(1)
    regs.close()
    print("case %3d ok: P %-18s n %-12s %s %s %s iters %2d (cut %2d) burn %2d thin %d tune %2d" % (
        case, Ps, ns, "voigt" if voigt else "gauss", "sd" if sd_free else "noise", "lsf" if use_lsf else "   ", iters, cut,
        burn, thin, tune), flush=True)
print("%d cases, all chains = replay step for step; worst trace deviation %.3g of the tolerance (1e-9 + 1e-8 |x|); %.0f s" % (
    cases, worst, time.perf_counter() - t_start))
