"""GPU: K4 (detection ratio) against the oracle / the reference's goldens, and VPspectrum end to end."""
import ast
import os

import numpy as np
import pytest

from oracle import detect as OD
from vamp_b200 import physics, synth
from vamp_b200 import vpspectrum as VS

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "detect.npz"), allow_pickle=False)
NAMES = [str(n) for n in G["names"]]
SETTINGS = [dict(ast.literal_eval(str(s))) for s in G["settings"]]


def _close(a, b):
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    assert np.array_equal(a[~fin], b[~fin])            # -inf end pixels
    np.testing.assert_allclose(a[fin], b[fin], rtol=1e-11, atol=1e-13)


def test_detection_ratio_matches_oracle_batched(ctx):
    """All golden sightlines (ragged: 90 ... 1001 px, odd and even lengths) in ONE launch."""
    for kw in SETTINGS:
        smin, smax = kw.get("std_min", 2), kw.get("std_max", 11)
        det = VS.detection_ratio_batch([G[n + "/wavelength"] for n in NAMES], [G[n + "/flux"] for n in NAMES],
                                       [G[n + "/noise"] for n in NAMES], smin, smax, ctx=ctx)
        for n, d in zip(NAMES, det):
            _close(d, OD.detection_ratio(G[n + "/wavelength"], G[n + "/flux"], G[n + "/noise"], smin, smax))


@pytest.mark.parametrize("name", NAMES)
def test_regions_match_reference_goldens(ctx, name):
    sp = VS.VPspectrum(1215.0).set_spectrum(G[name + "/wavelength"], G[name + "/flux"], G[name + "/noise"])
    for j, kw in enumerate(SETTINGS):
        sp.compute_detection_regions(**kw)
        assert sp.region_pixels == G["%s/s%d/region_pixels" % (name, j)].tolist()
        assert sp.region_waves[0][0] == sp.wavelength_array[sp.region_pixels[0][0]]


def test_detection_edge_cases(ctx):
    rng = np.random.default_rng(3)
    cases = []
    for P in (3, 4, 31, 256, 257, 700):                 # shorter than the Gaussian support, chunk boundaries
        wave = 1000.0 + 0.05 * np.arange(P) + 0.001 * rng.uniform(size=P).cumsum()
        flux = 1.0 - 0.5 * np.exp(-0.5 * ((np.arange(P) - P / 2.0) / 3.0) ** 2) + rng.normal(0, 0.02, P)
        noise = np.full(P, 0.02)
        if P == 31:
            noise[:] = 0.0                               # zero noise: 1/sqrt(0) = inf, 0 * inf = NaN never wins
        cases.append((wave, flux, noise))
    det = VS.detection_ratio_batch([c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases], ctx=ctx)
    for (w, f, n), d in zip(cases, det):
        with np.errstate(all="ignore"):
            ref = OD.detection_ratio(w, f, n)
        fin = np.isfinite(ref)
        assert np.array_equal(np.isnan(d), np.isnan(ref))
        np.testing.assert_allclose(d[fin], ref[fin], rtol=1e-11, atol=1e-13)
        assert np.array_equal(d[np.isinf(ref)], ref[np.isinf(ref)])


def _sightline(P=900, seed=11, snr=60.0, centers=(150.0, 420.0, 470.0, 760.0)):
    rng = np.random.default_rng(seed)
    wave = 1215.0 + 0.0203 * np.arange(P)                 # ~5 km/s pixels
    nu = physics.Wave2freq(wave)
    truth = []
    tau = np.zeros(P)
    for c in centers:
        amp, sig_px = rng.uniform(0.8, 2.5), rng.uniform(4.0, 8.0)
        sig = abs(nu[int(c)] - nu[int(c) + 1]) * sig_px
        tau += amp * np.exp(-0.5 * ((nu - nu[int(c)]) / sig) ** 2)
        truth.append((amp, nu[int(c)], sig))
    noise = np.full(P, 1.0 / snr)
    flux = np.exp(-tau) + rng.normal(0, 1.0 / snr, P)
    return wave, flux, noise, np.array(truth), np.exp(-tau)


def test_fit_spectrum_end_to_end(ctx):
    wave, flux, noise, truth, clean = _sightline()
    sp = VS.VPspectrum(1215.0, voigt=False, chi_limit=1.5, convergence_attempts=3, iterations=1500, burn=300, thin=5,
                       restarts=3, chains_per_restart=16, window=3, nmax=5, seed=4).set_spectrum(wave, flux, noise)
    params = sp.fit_spectrum()
    assert len(sp.region_pixels) == 4 and sp.region_pixels[0][0] < 150 < sp.region_pixels[0][1]
    for key in ("b", "b_std", "N", "N_std", "EW", "centers", "region_numbers"):
        assert len(params[key]) == len(params["b"]) >= 4
    fm = sp.flux_model
    assert fm["total"].shape == flux.shape and np.all(fm["chi_squared"] < 1.5)
    assert np.max(np.abs(fm["total"] - clean)) < 0.08      # the model follows the noiseless spectrum
    # every true line is recovered: centre within half a pixel, N within 15 %
    N_true = physics.ColumnDensity(truth[:, 0], truth[:, 2])
    cen_true = physics.Freq2wave(truth[:, 1])
    for Nt, ct in zip(N_true, cen_true):
        k = int(np.argmin(np.abs(params["centers"] - ct)))
        assert abs(params["centers"][k] - ct) < 0.5 * 0.0203
        assert params["N"][k] == pytest.approx(Nt, rel=0.15)
        assert params["N_std"][k] > 0 and params["b_std"][k] > 0
        assert params["b"][k] == pytest.approx(physics.DopplerParameter(truth[:, 2], 1215.0)[np.argmin(np.abs(cen_true - ct))], rel=0.15)
    # per-region component fluxes and equivalent widths are consistent with the total
    for j, (s, e) in enumerate(sp.region_pixels):
        comp = fm["region_%d_flux" % j]
        np.testing.assert_allclose(np.prod(comp, axis=0), fm["total"][s:e], rtol=1e-9)
        assert fm["EW"][j] > 0


def test_fit_spectrum_lsf_convolved(ctx):
    """BASELINE config C3 in miniature: an LSF-convolved sightline (Gaussian LSF, FWHM 2 px, 7 taps) split into
    regions and fitted with the LSF inside the model - the recovered widths are the intrinsic ones."""
    lsf = synth.gaussian_lsf(2.0, 7)
    wave, flux0, noise, truth, clean = _sightline(P=700, seed=23, snr=80.0, centers=(140.0, 380.0, 560.0))
    h = lsf.size // 2
    conv = 1.0 + np.convolve(np.concatenate([np.zeros(h), clean - 1.0, np.zeros(h)]), lsf, "valid")
    rng = np.random.default_rng(5)
    flux = conv + rng.normal(0, 1.0 / 80.0, len(wave))
    sp = VS.VPspectrum(1215.0, voigt=False, chi_limit=1.5, convergence_attempts=2, iterations=1500, burn=300, thin=5,
                       restarts=3, chains_per_restart=16, window=2, nmax=4, seed=2, lsf=lsf).set_spectrum(wave, flux, noise)
    params = sp.fit_spectrum()
    assert np.all(sp.flux_model["chi_squared"] < 1.6)
    for amp, cen, sig in truth:
        k = int(np.argmin(np.abs(params["centers"] - physics.Freq2wave(cen))))
        assert abs(params["centers"][k] - physics.Freq2wave(cen)) < 0.5 * 0.0203
        # b from the intrinsic width: a fit without the LSF in the model would be broadened by ~(0.85 px)^2 in quadrature
        assert params["b"][k] == pytest.approx(physics.DopplerParameter(sig, 1215.0), rel=0.12)
