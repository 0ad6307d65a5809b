"""GPU: the adaptive-Metropolis step method (``step_kind = VAMP_STEP_ADAPTIVE``; north_star "Metropolis /
adaptive-Metropolis", the reference's ``mcmc_fit(step_method=mc.AdaptiveMetropolis)``, vamp_1.0/vpfits.py:361-389).

* step for step: fp64 chains against ``oracle/chain_replay.py:replay_chain_am`` (PyMC 2 AdaptiveMetropolis semantics
  restated: per-component joint proposals from the running covariance, delay / interval / greedy tallying) over the
  oracle density, same Philox stream - proposals, accepts, covariance updates, Cholesky factors, tallies;
* stationarity: pooled adaptive chains give the posterior the scalar sampler gives (same target density), and adapt
  to an acceptance rate a random-walk sampler is expected to reach.
"""
import numpy as np
import pytest

from oracle import chain_replay as CR
from vamp_b200 import _lib, engine, synth

pytestmark = pytest.mark.gpu


def _scaled_grid(regs, r=0):
    nu = np.asarray(regs.host_regions[r][0])
    return (nu - regs.nu0[r]) / regs.scale[r]


@pytest.mark.parametrize("voigt,sd_free,P,n,lsf", [(True, True, 128, 2, None), (False, True, 90, 1, None),
                                                   (True, False, 60, 1, None), (False, True, 200, 3, "lsf")])
def test_adaptive_chain_replay_fp64(ctx, voigt, sd_free, P, n, lsf):
    lsf = synth.gaussian_lsf(2.0, 7) if lsf else None
    reg = synth.make_region(P, n, 40.0, 1300 + P, voigt=voigt, lsf=lsf)
    regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])], lsf=lsf)
    B, iters, burn, thin, tune, delay, interval = 3, 260, 60, 5, 25, 50, 20
    uid = np.array([11, 2 ** 34 + 3, 987654321], dtype=np.uint64)
    seed = 0xA11CE
    cmax = n + 1
    s = engine.Sampler(regs, np.zeros(B, np.int32), np.full(B, n, np.int32), uid, voigt=voigt, sd_free=sd_free, dtype="f64",
                       seed=seed, burn=burn, thin=thin, tune_interval=tune, tune_throughout=True,
                       trace_cap=(iters - burn) // thin + 2, cmax=cmax, step_method="adaptive", am_delay=delay,
                       am_interval=interval)
    # start near the truth: the joint proposals are accepted from the start, so the greedy history fills
    start = np.zeros((B, s.ts))
    t = regs.to_scaled(reg["truth"][None], [0])[0]
    for b in range(B):
        start[b, : 4 * n] = (t * (1.0 + 0.01 * (b + 1))).reshape(-1)
    if sd_free:
        start[:, 4 * cmax] = 1.0 / 40.0
    s.init(start)
    th0 = s.get("theta").copy()
    NP = 4 if voigt else 3
    am0 = s.get("am_state")
    for c in range(n):       # C_0 = diag(|value| / 100) and its factor
        np.testing.assert_allclose([am0[0, c, 4 + i * (i + 1) // 2 + i] for i in range(NP)], np.abs(th0[0, 4 * c: 4 * c + NP]) / 100.0)
        np.testing.assert_allclose([am0[0, c, 14 + i * (i + 1) // 2 + i] for i in range(NP)],
                                   np.sqrt(np.abs(th0[0, 4 * c: 4 * c + NP]) / 100.0))
    s.run(130)
    s.run(iters - 130)       # resume mid-way: adaptive state persists across launches
    u = _scaled_grid(regs)
    th, tr, ntal, am = s.get("theta"), s.get("trace"), s.get("ntally"), s.get("am_state")
    acct = s.get("acc_total")
    for b in range(B):
        ref = CR.replay_chain_am(u, reg["flux"], reg["noise"], th0[b], n, voigt, sd_free, int(uid[b]), seed, iters, burn, thin,
                                 tune, True, cmax, delay=delay, interval=interval, lsf=lsf)
        assert ntal[b] == ref["ntally"] == len(ref["trace"])
        np.testing.assert_allclose(tr[b, : ntal[b]], ref["trace"], rtol=1e-7, atol=1e-9, err_msg="chain %d" % b)
        np.testing.assert_allclose(th[b], ref["theta"], rtol=1e-7, atol=1e-9)
        for c in range(n):
            Cd = np.zeros((NP, NP))
            Ld = np.zeros((NP, NP))
            for i in range(NP):
                for j in range(i + 1):
                    Cd[i, j] = Cd[j, i] = am[b, c, 4 + i * (i + 1) // 2 + j]
                    Ld[i, j] = am[b, c, 14 + i * (i + 1) // 2 + j]
            # (the covariance is a difference of second moments: traces equal to 1e-9 leave it equal to ~1e-6)
            np.testing.assert_allclose(Cd, ref["cov"][c], rtol=2e-5, atol=1e-9 * np.abs(ref["cov"][c]).max())
            np.testing.assert_allclose(Ld, ref["chol"][c], rtol=2e-5, atol=1e-6 * np.abs(ref["chol"][c]).max())
            np.testing.assert_allclose(am[b, c, :NP], ref["chain_mean"][c], rtol=1e-7)
            assert am[b, c, 38] > 0                         # the covariance was updated from the history
        assert ref["accepted"].sum() > 0


def test_adaptive_needs_its_state_and_single_lines(ctx):
    reg = synth.make_region(64, 1, 40.0, 5, voigt=True)
    regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])])
    s = engine.Sampler(regs, np.zeros(2, np.int32), np.ones(2, np.int32), np.arange(2, dtype=np.uint64), step_method="adaptive")
    s.init()
    s.t["am_state"] = None
    with pytest.raises(_lib.VampError, match="am_state"):
        s.run(5)
    with pytest.raises(ValueError):
        engine.Sampler(regs, np.zeros(2, np.int32), np.ones(2, np.int32), np.arange(2, dtype=np.uint64), step_method="nuts")


def test_adaptive_and_scalar_samplers_share_the_posterior(ctx):
    """Same target density: pooled fp32 chains of both step methods, started at the truth, agree on the posterior of a
    2-component Gaussian region; the adaptive chains mix at least as well (their joint proposals follow the
    amplitude - width correlation the scalar steps have to zig-zag along)."""
    reg = synth.make_region(160, 2, 40.0, 4321, voigt=False)
    regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])])
    B, iters, burn, thin = 256, 12000, 4000, 10
    out = {}
    for method in ("metropolis", "adaptive"):
        s = engine.Sampler(regs, np.zeros(B, np.int32), np.full(B, 2, np.int32), np.arange(B, dtype=np.uint64) + 100, voigt=False,
                           dtype="f32", seed=8, burn=burn, thin=thin, tune_interval=100, tune_throughout=True,
                           step_method=method, am_delay=500, am_interval=100)
        start = np.zeros((B, s.ts))
        start[:, :8] = regs.to_scaled(reg["truth"][None], [0])[0].reshape(-1)
        start[:, 8] = 1.0 / 40.0
        s.init(start)
        s.run(iters)
        mean, m2, nt = s.get("mean"), s.get("m2"), s.get("ntally")
        assert nt.min() == nt.max() == (iters - burn) // thin
        acc = s.get("acc_total").astype(np.float64)
        out[method] = dict(mean=mean.mean(axis=0), std=np.sqrt((m2 / nt[:, None]).mean(axis=0) + mean.var(axis=0)),
                           mc=mean.std(axis=0, ddof=1) / np.sqrt(B), rate=acc[:, 0].mean() / (iters - burn))
    a, m = out["adaptive"], out["metropolis"]
    free = [0, 1, 2, 4, 5, 6, 8]
    tol = 5.0 * np.sqrt(a["mc"] ** 2 + m["mc"] ** 2)[free] + 0.03 * m["std"][free]
    assert np.all(np.abs(a["mean"][free] - m["mean"][free]) <= tol), (a["mean"][free] - m["mean"][free]) / tol
    np.testing.assert_allclose(a["std"][free], m["std"][free], rtol=0.15)
    # an adapted 3-dimensional random walk accepts 15 - 50 % of its proposals (optimal ~ 0.3)
    assert 0.12 < a["rate"] < 0.55, a["rate"]
    # between-chain scatter of the means (a mixing measure): adaptive not worse than 1.5 x scalar
    assert np.all(a["mc"][free] <= 1.5 * m["mc"][free] + 1e-12)
