"""GPU: BASELINE config C3 - a SIMBA-like synthetic sightline seen in HI 1215, CIV 1548/1550 and MgII 2796/2803
(doublets with tied v / b / N), LSF-convolved (Gaussian, FWHM 2 px, 7 taps), 1000 pixels per ion, split into regions
and fitted through ``VPspectrum.fit_spectrum`` on one GPU; and the result harvest key by key against the
oracle's restatement of vamp_1.0/vpspectrum.py:351-426 on the same MAP parameters."""
import numpy as np
import pytest

from oracle import harvest as OH
from vamp_b200 import physics, synth
from vamp_b200 import vpspectrum as VS

pytestmark = pytest.mark.gpu

FIT = dict(voigt=True, chi_limit=1.5, convergence_attempts=2, iterations=2000, burn=300, thin=5, restarts=3,
           chains_per_restart=32, window=2, nmax=3, seed=3)


def fit_c3(sight, **over):
    kw = dict(FIT)
    kw.update(over)
    sp = VS.VPspectrum(sight["line"], doublet=sight["doublet"], lsf=sight["lsf"], **kw)
    sp.set_spectrum(sight["wave"], sight["flux"], sight["noise"])
    sp.fit_spectrum()
    return sp


@pytest.mark.parametrize("k", [0, 1, 2])
def test_c3_sightline_recovers_all_absorbers(ctx, k):
    sight = synth.config_c3()[k]
    sp = fit_c3(sight)
    params = sp.params
    truth = sight["truth"]
    assert len(sp.region_pixels) == 3, sp.region_pixels               # one (joint) region per absorber
    for (s, e), (px, _, _) in zip(sp.region_pixels, sight["systems"]):
        assert s < px < e and e - s <= 256
    assert [r.n for r in sp.regions] == [1, 1, 1]
    assert np.all(sp.flux_model["chi_squared"] < 1.5)
    # the model (before the LSF) follows the intrinsic spectrum within the LSF's smoothing
    assert np.max(np.abs(sp.flux_model["total"] - sight["clean_intrinsic"])) < 0.12
    cen_true = physics.Freq2wave(truth[:, 1])
    sig_true = truth[:, 3] / 2.3548200450309493
    dw = sight["wave"][1] - sight["wave"][0]
    for i in range(3):
        j = int(np.argmin(np.abs(params["centers"] - cen_true[i])))
        assert abs(params["centers"][j] - cen_true[i]) < 0.5 * dw                                       # v
        assert params["b"][j] == pytest.approx(physics.DopplerParameter(sig_true[i], sight["line"]), rel=0.15)   # b
        th = sp.regions[i].theta[0]
        assert th[0] * synth.voigt_peak_shape(th[2], th[3]) == pytest.approx(
            truth[i, 0] * synth.voigt_peak_shape(truth[i, 2], truth[i, 3]), rel=0.15)                  # tau_0 (N up to L/G)
        assert params["N"][j] > 0 and params["N_std"][j] > 0 and params["EW"][j] > 0


@pytest.mark.parametrize("mode", ["get_mcmc_err", "mcmc_cov"])
def test_harvest_key_by_key_against_oracle(ctx, mode):
    """``params`` / ``flux_model`` of the product against the oracle's restatement of the reference's
    harvest lines, fed with the SAME per-region MAP parameters and posterior moments (HI sightline: single lines, the
    only case the reference has)."""
    sight = synth.config_c3(lsf=False)[0]
    sp = VS.VPspectrum(sight["line"], mcmc_cov=(mode == "mcmc_cov"), get_mcmc_err=True, **FIT)
    sp.set_spectrum(sight["wave"], sight["flux"], sight["noise"])
    sp.fit_spectrum()
    fits = []
    for r in sp.regions:
        f = np.sqrt(r.ntally / (r.ntally - 1.0))
        sig_std = r.std[:, 3] / 2.3548200450309493
        cov = np.zeros((r.n, 3, 3))
        cov[:, 0, 0], cov[:, 1, 1], cov[:, 2, 2] = (r.std[:, 0] * f) ** 2, (sig_std * f) ** 2, (r.std[:, 1] * f) ** 2
        cov[:, 0, 1] = cov[:, 1, 0] = r.cov_aw
        fits.append({"theta": r.theta, "cov": cov, "std_a": r.std[:, 0], "std_s": sig_std})
    params, fm = OH.harvest(sp.wavelength_array, sp.frequency_array, sp.flux_array, sp.region_pixels, fits, sight["line"],
                            True, [r.best_chi_squared for r in sp.regions], mcmc_cov=(mode == "mcmc_cov"), get_mcmc_err=True)
    assert sorted(params) == sorted(sp.params)
    for key in params:
        np.testing.assert_allclose(sp.params[key], params[key], rtol=1e-9, atol=0, err_msg=key)
    assert sorted(fm) == sorted(sp.flux_model)
    for key in fm:
        if key in ("region_pixels", "difficult_fit"):
            assert fm[key] == sp.flux_model[key]
        else:
            # 'total' and the per-component fluxes come from the device Faddeeva (float64, |dH| <= 7e-15)
            np.testing.assert_allclose(sp.flux_model[key], fm[key], rtol=1e-9, atol=1e-12, err_msg=key)
