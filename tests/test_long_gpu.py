"""GPU: regions longer than one 256-pixel coordinate tile (the reference fits arrays of any length - 2000 pixels in
vamp_1.0/vpfits_intro.ipynb).  The long-region sampler kernels (vamp_launch_k2l_*: 257 .. 2048 pixels, one chain per
CTA, all tiles of the region in shared memory relative to the first tile's base) and the tile-aware CTA schedule of
the MAP kernel against the same oracles as the one-tile kernels."""
import numpy as np
import pytest
from scipy.optimize import fmin

from oracle import chain_replay as CR
from oracle import density as D
from vamp_b200 import _lib, engine, synth
from vamp_b200.fit_regions import fit_regions
from vamp_b200.vpfits import VPfit

pytestmark = pytest.mark.gpu


def _scaled_grid(regs, r=0):
    nu = np.asarray(regs.host_regions[r][0])
    return (nu - regs.nu0[r]) / regs.scale[r]


@pytest.mark.parametrize("voigt,sd_free,P,n,lsf", [
    (True, True, 300, 2, None),          # two tiles
    (True, False, 700, 3, None),         # four tiles, per-pixel noise
    (False, True, 1100, 2, "lsf"),       # eight-tile buffers, LSF
    (True, True, 2000, 2, None),         # the intro notebook's length
])
def test_long_chain_replay_fp64(ctx, voigt, sd_free, P, n, lsf):
    """fp64 chains of a long region reproduce the CPU replay (oracle density on the whole grid) step for step."""
    lsf = synth.gaussian_lsf(2.0, 7) if lsf else None
    reg = synth.make_region(P, n, 40.0, 1900 + P, voigt=voigt, lsf=lsf)
    regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])], lsf=lsf)
    B, iters, burn, thin, tune = 3, 60, 20, 5, 25
    uid = np.array([11, 2 ** 34 + 3, 987654321], dtype=np.uint64)
    seed = 0xABCDEF12345
    s = engine.Sampler(regs, np.zeros(B, np.int32), np.full(B, n, np.int32), uid, voigt=voigt, sd_free=sd_free,
                       dtype="f64", seed=seed, burn=burn, thin=thin, tune_interval=tune, tune_throughout=True,
                       trace_cap=(iters - burn) // thin + 2, cmax=n + 1)
    s.init()
    th0 = s.get("theta").copy()
    s.run(25)
    s.run(iters - 25)
    u = _scaled_grid(regs)
    th, tr, ntal, lnL = s.get("theta"), s.get("trace"), s.get("ntally"), s.get("lnL")
    for b in range(B):
        ref = CR.replay_chain(u, reg["flux"], reg["noise"], th0[b], n, voigt, sd_free, int(uid[b]), seed, iters, burn,
                              thin, tune, True, n + 1, lsf=lsf)
        assert ntal[b] == ref["ntally"]
        np.testing.assert_allclose(tr[b, : ntal[b]], ref["trace"], rtol=1e-8, atol=1e-9, err_msg="chain %d" % b)
        np.testing.assert_allclose(th[b], ref["theta"], rtol=1e-8, atol=1e-9)
        assert lnL[b] == pytest.approx(ref["lnL"], rel=1e-9, abs=1e-6)


def test_long_fp32_tracks_fp64(ctx):
    """The fp32 long-region kernel shares the RNG stream with the fp64 one: same accept / reject history at first."""
    reg = synth.make_region(900, 3, 50.0, 31, voigt=True)
    regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])])
    out = {}
    for dt in ("f32", "f64"):
        s = engine.Sampler(regs, np.zeros(8, np.int32), np.full(8, 3, np.int32), np.arange(8, dtype=np.uint64) + 70,
                           voigt=True, dtype=dt, seed=9, burn=0, thin=1, tune_interval=10, tune_throughout=True)
        s.init()
        s.run(12)
        out[dt] = s.get("theta")
    np.testing.assert_allclose(out["f32"], out["f64"], rtol=5e-3, atol=5e-3)


@pytest.mark.parametrize("dtype", ["f64", "mixed"])
def test_long_map_polish_against_scipy_fmin(ctx, dtype):
    """K3 on a 700-pixel region (CTA schedule, tile by tile; "mixed" falls back to it) = scipy fmin on the oracle density."""
    voigt, sd_free, P = True, True, 700
    reg = synth.make_region(P, 2, 50.0, 778, voigt=voigt)
    regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])])
    u = _scaled_grid(regs)
    n, cmax, S = 2, 2, 4
    rng = np.random.default_rng(5)
    truth_s = regs.to_scaled(reg["truth"][None], [0])[0]
    starts = np.zeros((S, 4 * cmax + 1))
    for i in range(S):
        starts[i, : 4 * n] = (truth_s * (1 + rng.uniform(-0.2, 0.2, truth_s.shape))).reshape(-1)
        starts[i, 4 * cmax] = 0.05
    s = engine.Sampler(regs, np.zeros(S, np.int32), np.full(S, n, np.int32), np.arange(S, dtype=np.uint64), voigt=voigt,
                       sd_free=sd_free, dtype="f64", cmax=cmax)
    res = s.map_polish(starts, np.zeros(S, np.int32), np.full(S, n, np.int32), maxiter=4000, xtol=1e-4, ftol=1e-3,
                       dtype=dtype)
    free = [4 * c + i for c in range(n) for i in range(4)] + [4 * cmax]

    def neg_logp(p, full):
        full = full.copy()
        full[free] = p
        lp = D.log_posterior(full[: 4 * n].reshape(n, 4), u, reg["flux"], voigt, sd=full[4 * cmax])
        if np.isfinite(lp):
            c_lo, c_hi, w_max = D.prior_bounds(u, voigt)
            lp += n * np.log(c_hi - c_lo) + n * 2 * np.log(w_max)
        return -lp if np.isfinite(lp) else np.inf

    th = res["theta"].cpu().numpy()
    for i in range(S):
        _, fopt, *_ = fmin(neg_logp, starts[i][free], args=(starts[i],), xtol=1e-4, ftol=1e-3, maxiter=4000, maxfun=10 ** 9,
                           full_output=True, disp=0)
        f_dev = -float(res["logp"][i])
        assert neg_logp(th[i][free], starts[i]) == pytest.approx(f_dev, rel=1e-9, abs=1e-7)
        assert f_dev == pytest.approx(fopt, abs=5e-3)
        comps = th[i][: 4 * n].reshape(n, 4)
        assert float(res["chi2"][i]) == pytest.approx(D.chisquared(reg["flux"], D.model_flux(u, comps, voigt), reg["noise"]), rel=1e-8)


def test_vpfit_fits_a_2000_pixel_array(ctx):
    """vpfits_intro.ipynb fits a 2000-pixel array with VPfit directly: three Voigt lines recovered, read-backs = oracle."""
    reg = synth.make_region(2000, 3, 50.0, region_id=77, voigt=True, ratio_range=(0.1, 0.5))
    fit = VPfit(seed=3, n_chains=32)
    fit.find_bic(reg["nu"], reg["flux"], 3, reg["noise"], len(reg["nu"]) - 9, voigt=True, iterations=3000, thin=15, burn=300)
    assert min(fit.red_chi_array) < 1.3
    ev = [fit.estimated_variables[i] for i in range(3)]
    got = np.sort([v["centroid"].value for v in ev])
    np.testing.assert_allclose(got, np.sort(reg["truth"][:, 1]), atol=1.0 * synth.DNU)
    th = np.array([[v["amplitude"].value, v["centroid"].value, v["L_fwhm"].value, v["G_fwhm"].value] for v in ev])
    np.testing.assert_allclose(fit.total.value, D.model_flux(reg["nu"], th, True), rtol=1e-6)
    lnL = D.normal_loglike(reg["flux"], fit.total.value, sd=fit._sd)
    assert fit.map.BIC == pytest.approx(D.bic(lnL, 13, 2000), rel=1e-6)


def test_mixed_lengths_do_not_change_short_regions(ctx):
    """A region set holding short and long regions: the short ones run the one-tile kernels in their own launches and
    give the records they give alone (Philox keyed by region uid); the long one lands on its absorbers."""
    short = [synth.make_region(P, n, 40.0, 500 + P, voigt=True) for P, n in ((100, 1), (180, 2), (250, 1))]
    long_ = synth.make_region(600, 2, 40.0, 1234, voigt=True)
    kw = dict(voigt=True, iterations=1500, burn=300, thin=15, restarts=2, chains_per_restart=8, window=2, nmax=3, seed=4)
    host_s = [(r["nu"], r["flux"], r["noise"]) for r in short]
    alone = fit_regions(host_s, uids=[1, 2, 3], n0=[1, 2, 1], **kw)
    mixed = fit_regions(host_s + [(long_["nu"], long_["flux"], long_["noise"])], uids=[1, 2, 3, 4], n0=[1, 2, 1, 2], **kw)
    for i in range(3):
        a, m = alone[i], mixed[i]
        assert a.n == m.n
        for n in a.bic:
            np.testing.assert_array_equal(a.bic[n], m.bic[n])
        np.testing.assert_array_equal(a.theta, m.theta)
    r = mixed[3]
    assert r.n >= 2
    got = np.sort(r.theta[:, 1])
    for c in long_["truth"][:, 1]:
        assert np.min(np.abs(got - c)) < 1.5 * synth.DNU


def test_limits_fail_loudly(ctx):
    too_long = synth.make_region(2100, 1, 40.0, 5, voigt=True)
    regs = engine.Regions(ctx, [(too_long["nu"], too_long["flux"], too_long["noise"])])
    with pytest.raises(_lib.VampError, match="2048"):
        engine.Sampler(regs, np.zeros(2, np.int32), np.ones(2, np.int32), np.arange(2, dtype=np.uint64), voigt=True)
    reg = synth.make_region(2000, 1, 40.0, 6, voigt=True)
    regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])])
    s = engine.Sampler(regs, np.zeros(2, np.int32), np.full(2, 30, np.int32), np.arange(2, dtype=np.uint64), voigt=True)
    s.init()
    with pytest.raises(_lib.VampError, match="shared memory"):
        s.run(5)
    with pytest.raises(_lib.VampError, match="adaptive"):
        engine.Sampler(regs, np.zeros(2, np.int32), np.ones(2, np.int32), np.arange(2, dtype=np.uint64), voigt=True,
                       step_method="adaptive")


def test_vpspectrum_fits_a_blended_complex_whole(ctx):
    """A 360-pixel blended complex (no continuum inside).  The default (``max_region_pixels=256``) has to cut it through
    absorption and says so; ``max_region_pixels=2048`` fits the detection region whole on the long-region kernels, as
    the reference does, and recovers every line."""
    import warnings
    from vamp_b200 import physics, vpspectrum as VS
    P, snr = 900, 60.0
    rng = np.random.default_rng(3)
    wave = 1215.0 + 0.0203 * np.arange(P)
    nu = physics.Wave2freq(wave)
    tau, truth = np.zeros(P), []
    for c in 300 + 45 * np.arange(7):
        amp, sig_px = rng.uniform(1.0, 2.0), rng.uniform(10.0, 13.0)
        sig = abs(nu[c] - nu[c + 1]) * sig_px
        tau += amp * np.exp(-0.5 * ((nu - nu[c]) / sig) ** 2)
        truth.append((amp, nu[c], sig))
    truth = np.array(truth)
    noise = np.full(P, 1.0 / snr)
    flux = np.exp(-tau) + rng.normal(0, 1.0 / snr, P)
    kw = dict(voigt=False, chi_limit=1.5, convergence_attempts=2, iterations=3000, burn=300, thin=15, restarts=3,
              chains_per_restart=16, window=2, nmax=4, seed=6)
    sp = VS.VPspectrum(1215.0, **kw).set_spectrum(wave, flux, noise)
    sp.compute_detection_regions(min_region_width=2)
    assert len(sp.region_pixels) == 1 and sp.region_pixels[0][1] - sp.region_pixels[0][0] > 256
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        sp.split_long_regions()
    assert len(sp.region_pixels) >= 2 and len(sp.cuts_through_absorption) >= 1 and len(w) >= 1

    sp = VS.VPspectrum(1215.0, max_region_pixels=2048, **kw).set_spectrum(wave, flux, noise)
    params = sp.fit_spectrum()
    assert len(sp.region_pixels) == 1 and "cuts_through_absorption" not in sp.flux_model
    assert sp.flux_model["chi_squared"][0] < 1.5
    assert np.max(np.abs(sp.flux_model["total"] - np.exp(-tau))) < 0.08
    cen_true = physics.Freq2wave(truth[:, 1])
    N_true = physics.ColumnDensity(truth[:, 0], truth[:, 2])
    assert len(params["centers"]) >= 7
    for Nt, ct in zip(N_true, cen_true):
        k = int(np.argmin(np.abs(params["centers"] - ct)))
        assert abs(params["centers"][k] - ct) < 1.0 * 0.0203
        assert params["N"][k] == pytest.approx(Nt, rel=0.2)


def test_long_map_with_lsf_reports_the_oracle_values(ctx):
    """K3 on a 900-pixel LSF-convolved region: the reported log-posterior, lnL and chi^2 at the returned optimum are the
    oracle's (LSF staging buffer sized by the region, tiles streamed from global memory), and the polish improves on
    its start."""
    lsf = synth.gaussian_lsf(2.0, 7)
    reg = synth.make_region(900, 2, 50.0, 4321, voigt=True, lsf=lsf)
    regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])], lsf=lsf)
    u = _scaled_grid(regs)
    n, cmax, S = 2, 2, 3
    rng = np.random.default_rng(8)
    truth_s = regs.to_scaled(reg["truth"][None], [0])[0]
    starts = np.zeros((S, 4 * cmax + 1))
    for i in range(S):
        starts[i, : 4 * n] = (truth_s * (1 + rng.uniform(-0.15, 0.15, truth_s.shape))).reshape(-1)
        starts[i, 4 * cmax] = 0.05
    s = engine.Sampler(regs, np.zeros(S, np.int32), np.full(S, n, np.int32), np.arange(S, dtype=np.uint64), voigt=True,
                       sd_free=True, dtype="f64", cmax=cmax)
    res = s.map_polish(starts, np.zeros(S, np.int32), np.full(S, n, np.int32), maxiter=3000, xtol=1e-4, ftol=1e-3, dtype="f64")
    th = res["theta"].cpu().numpy()
    c_lo, c_hi, w_max = D.prior_bounds(u, True)
    const = n * np.log(c_hi - c_lo) + n * 2 * np.log(w_max)
    for i in range(S):
        comps = th[i][: 4 * n].reshape(n, 4)
        lp = D.log_posterior(comps, u, reg["flux"], True, sd=th[i][4 * cmax], lsf=lsf) + const
        assert float(res["logp"][i]) == pytest.approx(lp, rel=1e-9, abs=1e-7)
        model = D.model_flux(u, comps, True, lsf=lsf)
        assert float(res["lnL"][i]) == pytest.approx(D.normal_loglike(reg["flux"], model, sd=th[i][4 * cmax]), rel=1e-9)
        assert float(res["chi2"][i]) == pytest.approx(D.chisquared(reg["flux"], model, reg["noise"]), rel=1e-8)
        lp0 = D.log_posterior(starts[i][: 4 * n].reshape(n, 4), u, reg["flux"], True, sd=0.05, lsf=lsf) + const
        assert lp > lp0
