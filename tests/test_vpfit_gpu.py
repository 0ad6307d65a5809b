"""GPU: the reference-facing API (VPfit / VPregion) end to end on synthetic absorbers with
known truth, read-backs checked against the oracle."""
import numpy as np
import pytest

from oracle import density as D
from vamp_b200 import synth
from vamp_b200.vpfits import VPfit
from vamp_b200.vpregion import VPregion

pytestmark = pytest.mark.gpu


def _sorted_by_centroid(th):
    return th[np.argsort(th[:, 1])]


def test_find_bic_c1(ctx):
    """BASELINE config 1: one Voigt component, ~200 px."""
    reg = synth.make_region(200, 1, 50.0, region_id=1, voigt=True, ratio_range=(0.1, 0.5))
    fit = VPfit(seed=1, n_chains=96)
    freedom = len(reg["nu"]) - 3
    fit.find_bic(reg["nu"], reg["flux"], 1, reg["noise"], freedom, voigt=True, iterations=3000, thin=15, burn=300)
    assert len(fit.bic_array) == 3 and len(fit.red_chi_array) == 3
    assert min(fit.red_chi_array) < 1.5
    v = fit.estimated_variables[0]
    t = reg["truth"][0]
    # amplitude_L and L_fwhm are degenerate for L << G (only their product is constrained):
    # compare the fitted optical-depth profile with the true one instead of the raw amplitudes
    assert v["centroid"].value == pytest.approx(t[1], abs=1.0 * synth.DNU)
    tau_true = D.voigt_tau(reg["nu"], t[1], t[0], t[2], t[3])
    tau_fit = fit.estimated_profiles[0].value
    assert np.abs(tau_fit - tau_true).max() < 0.15 * tau_true.max()
    # read-backs are consistent with the oracle at the fitted values
    th = np.array([[v["amplitude"].value, v["centroid"].value, v["L_fwhm"].value, v["G_fwhm"].value]])
    total = fit.total.value
    np.testing.assert_allclose(total, D.model_flux(reg["nu"], th, True), rtol=1e-6)
    np.testing.assert_allclose(fit.estimated_profiles[0].value, D.component_taus(reg["nu"], th, True)[0], rtol=1e-8, atol=1e-14)
    lnL = D.normal_loglike(reg["flux"], total, sd=fit._sd)
    assert fit.map.BIC == pytest.approx(D.bic(lnL, 5, 200), rel=1e-6)
    assert fit.map.AIC == pytest.approx(D.aic(lnL, 5), rel=1e-6)
    assert min(fit.red_chi_array) == pytest.approx(D.reduced_chisquared(reg["flux"], total, reg["noise"], freedom), rel=1e-6)
    # traces / stats / covariance of the winning chain
    tr_a = fit.mcmc.trace("xexp_0")[:]
    tr_g = fit.mcmc.trace("est_G_0")[:]
    tr_c = fit.mcmc.trace(fit.estimated_variables[0]["centroid"])[:]
    assert len(tr_a) == (3000 - 300) // 15 == len(tr_c)
    st = fit.mcmc.stats()
    assert st["xexp_0"]["standard deviation"] == pytest.approx(tr_a.std(), rel=1e-6)
    assert st["est_centroid_0"]["mean"] == pytest.approx(tr_c.mean(), rel=1e-12)
    assert "sd" in st
    cov = fit.chain_covariance(1, voigt=True)
    ref = np.cov(np.array((tr_a, VPfit.GaussianWidth(tr_g), tr_c)))
    np.testing.assert_allclose(cov[0], ref, rtol=1e-5, atol=1e-12 * np.abs(ref).max())
    assert np.isfinite(fit.mcmc.DIC) and np.isfinite(fit.mcmc.BPIC)
    assert fit.mcmc.trace("component_0")[:].shape == (len(tr_a), 200)


def test_notebook_flow_gaussian(ctx):
    """vpfits_intro.ipynb flow: initialise_model -> map_estimate -> mcmc_fit -> read-backs."""
    reg = synth.make_region(150, 2, 50.0, 31, voigt=False)
    fit = VPfit(seed=3, n_chains=48)
    fit.initialise_model(reg["nu"], reg["flux"], 2, voigt=False)
    assert fit.estimated_variables[0]["amplitude"].value == 0.5          # xexp initial value (vpfits.py:240)
    fit.map_estimate(iterations=3000)
    bic0 = fit.map.BIC
    fit.mcmc_fit(iterations=3000, burnin=500, thinning=10)
    assert hasattr(fit, "fit_time")
    tr = fit.mcmc.trace("est_sigma_0")[:]
    assert len(tr) == 250
    fit.map_estimate(iterations=3000)
    assert fit.map.BIC <= bic0 + 1.0
    got = _sorted_by_centroid(np.array([[fit.estimated_variables[c][k].value for k in ("amplitude", "centroid", "sigma")] for c in range(2)]))
    want = _sorted_by_centroid(reg["truth"][:, :3])
    np.testing.assert_allclose(got[:, 0], want[:, 0], rtol=0.2)
    np.testing.assert_allclose(got[:, 1], want[:, 1], atol=1.0 * synth.DNU)
    np.testing.assert_allclose(got[:, 2], want[:, 2], rtol=0.25)
    chi = VPfit.ReducedChisquared(reg["flux"], fit.total.value, reg["noise"], 150 - 6)
    assert chi < 1.6
    # the notebook's convergence cells (17, 18): mc.geweke / mc.raftery_lewis per variable, PyMC 2 names
    rl = fit.raftery_lewis(q=0.995, r=0.05)
    assert sorted(rl) == sorted(["xexp_0", "est_centroid_0", "est_sigma_0", "xexp_1", "est_centroid_1", "est_sigma_1", "sd"])
    assert all(v[0] == 8 and v[1] >= 1 and v[4] >= 1 for v in rl.values())     # nmin = 8 as printed in cell 18
    gz = fit.geweke(intervals=20)
    assert gz["est_centroid_0"].shape == (20, 2) and np.all(np.isfinite(gz["est_centroid_0"]))
    # a second run on the same sampler continues with its own random stream (not a replay of the first)
    fit.mcmc_fit(iterations=3000, burnin=500, thinning=10)
    tr2 = fit.mcmc.trace("est_sigma_0")[:]
    assert len(tr2) == 250 and not np.array_equal(tr2, tr)
    assert fit._sampler._runs == 2 and fit._sampler.seed != fit._seed


def test_fixed_noise_mode(ctx):
    reg = synth.make_region(100, 1, 40.0, 77, voigt=True, ratio_range=(0.1, 0.5))
    fit = VPfit(noise=reg["noise"], seed=5, n_chains=48)
    fit.find_bic(reg["nu"], reg["flux"], 1, reg["noise"], 97, voigt=True, iterations=2000, burn=300)
    assert "sd" not in fit.mcmc.stats()
    total = fit.total.value
    lnL = D.noise_map_loglike(reg["flux"], reg["noise"], total)
    assert fit.map.BIC == pytest.approx(D.bic(lnL, 4, 100), rel=1e-6)
    assert min(fit.red_chi_array) < 1.6


def test_region_fit_selects_component_count(ctx, capsys):
    """Two well separated Gaussian lines: the BIC ladder must stop at n = 2 (window replay)."""
    reg = synth.make_region(200, 2, 50.0, 2024, voigt=False)
    region = VPregion(reg["nu"], reg["flux"], reg["noise"], voigt=False, chi_limit=1.5, seed=11, n_chains=96)
    # noise minima make the reference's estimate_n start high at this S/N (vpregion.py:33);
    # start the ladder at 1 to exercise the whole acceptance rule
    region.n = 1
    region.set_freedom()
    assert region.freedom == 197
    region.region_fit(verbose=True, iterations=3000, burn=300)
    out = capsys.readouterr().out
    assert "Setting initial number of lines to: 1" in out
    assert len(region.fit.estimated_profiles) == 2
    region.set_freedom()
    chi = region.fit.ReducedChisquared(reg["flux"], region.fit.total.value, reg["noise"], region.freedom)
    assert chi < 1.5
    got = _sorted_by_centroid(np.array([[region.fit.estimated_variables[c][k].value for k in ("amplitude", "centroid", "sigma")] for c in range(2)]))
    want = _sorted_by_centroid(reg["truth"][:, :3])
    np.testing.assert_allclose(got[:, 1], want[:, 1], atol=1.0 * synth.DNU)


def test_simba_fixture_regions(ctx):
    """The reference's own HI fixture regions (simba_H1215.h5, pixel bounds of
    simba_spec_demo.ipynb cell 23): fits reach the reference's chi^2 acceptance level."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "simba.npz"))
    for j in range(3):
        nu, flux, noise = (g["simba_H1215/r%d/%s" % (j, k)] for k in ("nu", "flux", "noise"))
        region = VPregion(nu, flux, noise, voigt=False, chi_limit=1.5, seed=j, n_chains=96)
        assert region.n == int(g["simba_H1215/r%d/n0" % j]) and region.freedom == int(g["simba_H1215/r%d/freedom" % j])
        region.region_fit(verbose=False)
        region.set_freedom()
        chi = region.fit.ReducedChisquared(flux, region.fit.total.value, noise, region.freedom)
        assert chi < 2.5, (j, chi, region.n)
        assert 1 <= len(region.fit.estimated_profiles) <= 4     # notebook ladder ended at n = 2, 1, 1


def test_fit_regions_polish_schedules_agree(ctx):
    """The many-region polish schedule (one warp per start, fp32 search stage, float64 finish) selects the same
    component counts and the same optimum as the float64 CTA schedule."""
    from vamp_b200.fit_regions import fit_regions
    regs = [synth.make_region(P, n, 60.0, 3100 + i, voigt=False) for i, (P, n) in enumerate([(96, 1), (160, 2), (64, 1), (200, 2)])]
    host = [(r["nu"], r["flux"], r["noise"]) for r in regs]
    kw = dict(uids=[50, 51, 52, 53], voigt=False, iterations=1200, burn=300, thin=5, restarts=3, chains_per_restart=16,
              window=2, nmax=4, n0=[1, 1, 1, 1], seed=9, distributed=False)
    a = fit_regions(host, polish_dtype="f64", **kw)
    b = fit_regions(host, polish_dtype="mixed", **kw)
    assert [r.n for r in a] == [r.n for r in b]
    for ra, rb in zip(a, b):
        assert rb.lnL == pytest.approx(ra.lnL, abs=2e-2)
        np.testing.assert_allclose(rb.theta, ra.theta, rtol=2e-3)
        assert rb.chi2 == pytest.approx(ra.chi2, rel=1e-3)


def test_gelman_rubin_from_device_accumulators(ctx):
    """R-hat over the ~100 chains of a well-determined one-line fit is close to 1 for every sampled variable."""
    reg = synth.make_region(120, 1, 60.0, 515, voigt=False)
    fit = VPfit(seed=3, n_chains=96)
    fit.initialise_model(reg["nu"], reg["flux"], 1, voigt=False)
    fit.mcmc_fit(iterations=6000, burnin=3000, thinning=5)
    r = fit.gelman_rubin()
    assert set(r) == {"xexp_0", "est_centroid_0", "est_sigma_0", "sd"}
    assert all(0.98 < v < 1.2 for v in r.values()), r
