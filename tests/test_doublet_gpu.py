"""GPU: doublets with tied velocity / width / column density (BASELINE config C3: CIV 1548/1550, MgII 2796/2803).

Every component of a multiplet region carries a tied second line (``vamp_regions_set_multiplet``).  The reference
fits single lines only, so the pins are: (i) tests/golden/doublet.npz - sums of the reference's OWN profile-function
calls, one per line; (ii) the oracle density with ``multiplet`` (checked against (i) on the CPU); (iii) the fp64
chains step for step against the CPU replay; (iv) K3 against ``scipy.optimize.fmin`` on the oracle density; (v) a
full fit recovering a known CIV doublet.
"""
import os

import numpy as np
import pytest
from scipy.optimize import fmin

from oracle import chain_replay as CR
from oracle import density as D
from vamp_b200 import engine, physics, synth
from vamp_b200.fit_regions import fit_regions

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases():
    g = np.load(os.path.join(GOLDEN, "doublet.npz"))
    return g, [str(n) for n in g["names"]]


@pytest.mark.parametrize("dtype,tol_flux,tol_lnL", [("f64", 1e-6, 1e-6), ("f32", 1e-4, None)])
def test_k1_doublet_flux_and_lnL_match_reference_lines(ctx, dtype, tol_flux, tol_lnL):
    g, names = _cases()
    for name in names:
        nu, flux, noise = g[name + "/nu"], g[name + "/flux"], g[name + "/noise"]
        comps, voigt, mult = g[name + "/comps"], bool(g[name + "/voigt"]), tuple(g[name + "/multiplet"])
        lsf = g[name + "/lsf"]
        lsf = None if lsf.size == 1 else lsf
        regs = engine.Regions(ctx, [(nu, flux, noise)], lsf=lsf, multiplet=mult)
        want = D.lsf_convolve(g[name + "/total"], lsf)              # reference lines (+ this repo's LSF)
        for fn in (engine.loglike, engine.loglike_host):
            lnL, chi2, m = fn(regs, comps[None], [0], sd=[0.03], voigt=voigt, dtype=dtype, want_flux=True)
            np.testing.assert_allclose(m[0, : len(nu)], want, rtol=tol_flux, atol=tol_flux * 1e-3, err_msg=name)
            if tol_lnL:
                assert lnL[0] == pytest.approx(D.normal_loglike(flux, want, sd=0.03), rel=tol_lnL), name
                assert chi2[0] == pytest.approx(D.chisquared(flux, want, noise), rel=tol_lnL), name
        # the same region without the table is the single-line model
        single = engine.Regions(ctx, [(nu, flux, noise)], lsf=lsf)
        _, _, m1 = engine.loglike(single, comps[None], [0], sd=[0.03], voigt=voigt, dtype=dtype, want_flux=True)
        np.testing.assert_allclose(m1[0, : len(nu)], D.model_flux(nu, comps, voigt, lsf), rtol=tol_flux, atol=tol_flux * 1e-3)
        regs.set_multiplet(None)                                    # clearing the table restores it
        _, _, m2 = engine.loglike(regs, comps[None], [0], sd=[0.03], voigt=voigt, dtype=dtype, want_flux=True)
        np.testing.assert_array_equal(m1, m2)


def test_k1_doublet_long_region_and_mixed_table(ctx):
    """The CTA schedule of K1 (regions > 256 px) and a table where only some regions are doublets."""
    rng = np.random.default_rng(3)
    fr, ar, lr = physics.doublet_ratios("CIV")
    P = 3000
    nu = 1.9e15 + np.arange(P) * 3.2e10
    comps = np.array([[1.2, nu[2200], 2 * 3.2e10, 9 * 3.2e10], [0.6, nu[1500], 1 * 3.2e10, 5 * 3.2e10]])
    clean = D.model_flux(nu, comps, True, multiplet=(fr, ar, lr))
    flux = clean + rng.normal(0, 0.02, P)
    short = synth.make_doublet_region(256, 2, 50.0, 41, "CIV", True)
    regs = engine.Regions(ctx, [(nu, flux, np.full(P, 0.02)), (short["nu"], short["flux"], short["noise"]),
                                (short["nu"], short["flux"], short["noise"])], multiplet=[(fr, ar, lr), None, (fr, ar, lr)])
    th = np.zeros((3, 2, 4))
    th[0], th[1], th[2] = comps, short["truth"], short["truth"]
    lnL, chi2, m = engine.loglike(regs, th, [0, 1, 2], sd=[0.02] * 3, voigt=True, dtype="f64", want_flux=True)
    np.testing.assert_allclose(m[0], clean, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(m[1, :256], D.model_flux(short["nu"], short["truth"], True), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(m[2, :256], short["clean"], rtol=1e-6, atol=1e-9)


def _scaled(regs, mult, r=0):
    nu = np.asarray(regs.host_regions[r][0])
    u = (nu - regs.nu0[r]) / regs.scale[r]
    fr, ar, lr = mult
    return u, (fr, ar, lr, regs.nu0[r] * (fr - 1.0) / regs.scale[r])


@pytest.mark.parametrize("voigt,sd_free,doublet,n,lsf", [
    (True, True, "CIV", 2, None), (False, True, "CIV", 1, None), (True, False, "MgII", 1, None), (True, True, "CIV", 1, "lsf"),
])
def test_chain_replay_fp64_doublet(ctx, voigt, sd_free, doublet, n, lsf):
    """fp64 chains on a doublet region reproduce the CPU replay over the oracle's doublet density step for step."""
    lsf = synth.gaussian_lsf(2.0, 7) if lsf else None
    reg = synth.make_doublet_region(256 if doublet == "MgII" else 200, n, 40.0, 700 + n, doublet, voigt, lsf=lsf)
    regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])], lsf=lsf, multiplet=reg["multiplet"])
    P = len(reg["nu"])
    B, iters, burn, thin, tune = 3, 90, 30, 5, 20
    uid = np.array([11, 2 ** 34 + 1, 99], dtype=np.uint64)
    seed = 0xC1C1C1
    s = engine.Sampler(regs, np.zeros(B, np.int32), np.full(B, n, np.int32), uid, voigt=voigt, sd_free=sd_free, dtype="f64",
                       seed=seed, burn=burn, thin=thin, tune_interval=tune, tune_throughout=True,
                       trace_cap=(iters - burn) // thin + 2, cmax=n + 1)
    s.init()
    th0 = s.get("theta").copy()
    s.run(40)
    s.run(iters - 40)
    u, mult_s = _scaled(regs, reg["multiplet"])
    th, tr, ntal, lnL = s.get("theta"), s.get("trace"), s.get("ntally"), s.get("lnL")
    for b in range(B):
        ref = CR.replay_chain(u, reg["flux"], reg["noise"], th0[b], n, voigt, sd_free, int(uid[b]), seed, iters, burn, thin,
                              tune, True, n + 1, lsf=lsf, multiplet=mult_s)
        assert ntal[b] == ref["ntally"]
        np.testing.assert_allclose(tr[b, : ntal[b]], ref["trace"], rtol=1e-8, atol=1e-10, err_msg="chain %d" % b)
        np.testing.assert_allclose(th[b], ref["theta"], rtol=1e-8, atol=1e-10)
        assert lnL[b] == pytest.approx(ref["lnL"], rel=1e-9, abs=1e-7)
    # and the work counters count LINE evaluations: two per in-bounds centroid / width step
    w = s.work()
    assert w["shape_pix"] % (2 * P) == 0 and w["shape_pix"] > 0


@pytest.mark.parametrize("polish_dtype", ["f64", "mixed"])
def test_map_doublet_vs_fmin(ctx, polish_dtype):
    """K3 (both schedules) on a doublet region against scipy.optimize.fmin on the oracle's doublet density."""
    reg = synth.make_doublet_region(220, 1, 60.0, 801, "CIV", True)
    regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])], multiplet=reg["multiplet"])
    u, mult_s = _scaled(regs, reg["multiplet"])
    n, cmax = 1, 1
    start = regs.to_scaled(reg["truth"][None] * np.array([1.15, 1.0, 1.3, 0.85]), [0])[0]
    start[0, 1] += 0.8
    th0 = engine.pack_theta(start[None], [0.05], cmax)
    s = engine.Sampler(regs, np.zeros(1, np.int32), np.full(1, n, np.int32), np.array([5], dtype=np.uint64), voigt=True,
                       sd_free=True, dtype="f32", seed=1, cmax=cmax)
    torch = engine._torch()
    dev = regs.ctx.torch_device
    r = s._map(torch.from_numpy(th0).to(dev), torch.zeros(1, dtype=torch.int32, device=dev),
               torch.full((1,), n, dtype=torch.int32, device=dev), n, 3000, 0, 1e-4, 1e-3, polish_dtype)
    th = r["theta"].cpu().numpy()[0]

    def neg_logp(x):
        comps = x[:4].reshape(1, 4)
        P = len(u)
        if comps[0, 0] <= 0 or not (0 <= comps[0, 1] <= P - 1) or comps[0, 2] < 0 or comps[0, 3] <= 0 or not (0 < x[4] <= 1):
            return np.inf
        m = D.model_flux(u, comps, True, multiplet=mult_s)
        return -(D.normal_loglike(reg["flux"], m, sd=x[4]) + np.log(comps[0, 0]) - comps[0, 0])

    xo = fmin(neg_logp, th0[0], xtol=1e-4, ftol=1e-3, maxiter=3000, disp=False)
    assert -float(r["logp"][0]) == pytest.approx(neg_logp(xo), abs=5e-3)
    # lnL / chi^2 / BIC reported at the device optimum equal the oracle's at that point
    m = D.model_flux(u, th[:4].reshape(1, 4), True, multiplet=mult_s)
    lnL = D.normal_loglike(reg["flux"], m, sd=th[4])
    assert float(r["lnL"][0]) == pytest.approx(lnL, rel=1e-9)
    assert float(r["chi2"][0]) == pytest.approx(D.chisquared(reg["flux"], m, reg["noise"]), rel=1e-9)
    assert float(r["bic"][0]) == pytest.approx(5 * np.log(len(u)) - 2 * lnL, rel=1e-9)


def test_fit_regions_recovers_civ_doublet(ctx):
    """Config C3 in miniature: LSF-convolved CIV doublet regions (and a single-line region in the same batch) through
    the public API; the component count and the first line's parameters come back."""
    lsf = synth.gaussian_lsf(2.0, 7)
    dbl = [synth.make_doublet_region(256, n, 60.0, 900 + i, "CIV", True, lsf=lsf) for i, n in enumerate((1, 2, 2))]
    sgl = synth.make_region(128, 1, 60.0, 950, voigt=True, lsf=lsf)
    host = [(r["nu"], r["flux"], r["noise"]) for r in dbl] + [(sgl["nu"], sgl["flux"], sgl["noise"])]
    mult = [r["multiplet"] for r in dbl] + [None]
    res = fit_regions(host, voigt=True, iterations=3000, burn=300, thin=15, restarts=3, chains_per_restart=32, window=3,
                      n0=[1, 1, 1, 1], nmax=4, lsf=lsf, multiplet=mult, seed=4, distributed=False)
    for r, reg in zip(res, dbl + [sgl]):
        truth = reg["truth"][np.argsort(reg["truth"][:, 1])]
        assert r.n == len(truth), (r.n, len(truth), r.bic)
        fit = r.theta[np.argsort(r.theta[:, 1])]
        dnu = reg["nu"][1] - reg["nu"][0]
        np.testing.assert_allclose(fit[:, 1], truth[:, 1], atol=0.6 * dnu)                 # centroids within 0.6 px
        # central optical depth of the first line within 20 %
        t_fit = fit[:, 0] * synth.voigt_peak_shape(fit[:, 2], fit[:, 3])
        t_tru = truth[:, 0] * synth.voigt_peak_shape(truth[:, 2], truth[:, 3])
        np.testing.assert_allclose(t_fit, t_tru, rtol=0.2)
        np.testing.assert_allclose(fit[:, 3], truth[:, 3], rtol=0.2)                        # Doppler width within 20 %
    # fitting the doublet regions WITHOUT the table needs twice the components (each line on its own)
    res1 = fit_regions(host[:1], voigt=True, iterations=3000, burn=300, thin=15, restarts=3, chains_per_restart=32, window=3,
                       n0=[1], nmax=4, lsf=lsf, seed=4, distributed=False)
    assert res1[0].n == 2


def _civ_sightline(P=1000, seed=5, snr=70.0, systems=((200.0, 1.6, 5.0), (600.0, 1.1, 7.0)), lsf=None):
    """A CIV sightline: ~5 km/s pixels around 1548 A, each system = (pixel of line 1, tau0, sigma in px)."""
    rng = np.random.default_rng(seed)
    (l1, f1, g1), (l2, f2, g2) = physics.DOUBLETS["CIV"]
    wave = l1 * (1.0 - 0.004) + (l1 * 5.0e3 / 2.98e8) * np.arange(P)
    nu = physics.Wave2freq(wave)
    fr, ar, lr = physics.doublet_ratios("CIV")
    truth, tau = [], np.zeros(P)
    for px, tau0, sig_px in systems:
        c = nu[int(px)]
        G = 2.3548200450309493 * sig_px * abs(nu[int(px)] - nu[int(px) + 1])
        L = 0.05 * G
        A = tau0 / synth.voigt_peak_shape(L, G)
        tau += D.voigt_tau(nu, c, A, L, G) + D.voigt_tau(nu, fr * c, ar * A, lr * L, fr * G)
        truth.append((A, c, L, G))
    clean = np.exp(-tau)
    if lsf is not None:
        h = lsf.size // 2
        clean = 1.0 + np.convolve(np.concatenate([np.zeros(h), clean - 1.0, np.zeros(h)]), lsf, "valid")
    return wave, clean + rng.normal(0, 1.0 / snr, P), np.full(P, 1.0 / snr), np.array(truth), clean


def test_fit_spectrum_civ_doublet_sightline(ctx):
    """BASELINE config C3: an LSF-convolved sightline with CIV doublets - detection, pairing of the two lines' regions,
    the tied fit and the harvest of the first line's parameters."""
    from vamp_b200 import vpspectrum as VS
    lsf = synth.gaussian_lsf(2.0, 7)
    wave, flux, noise, truth, clean = _civ_sightline(lsf=lsf)
    sp = VS.VPspectrum(1548.204, voigt=True, doublet="CIV", chi_limit=1.5, convergence_attempts=2, iterations=2000, burn=300,
                       thin=5, restarts=3, chains_per_restart=32, window=2, nmax=3, seed=3, lsf=lsf).set_spectrum(wave, flux, noise)
    sp.compute_detection_regions(min_region_width=2)
    assert len(sp.region_pixels) == 4                       # two lines x two systems
    params = sp.fit_spectrum()
    assert len(sp.region_pixels) == 2                       # paired: one joint region per system
    for (s, e), px in zip(sp.region_pixels, (200, 600)):
        assert s < px < px + 99 < e and e - s <= 256
    assert [r.n for r in sp.regions] == [1, 1]
    assert np.all(sp.flux_model["chi_squared"] < 1.5)
    # the model (both lines of both systems, before the LSF) follows the unconvolved truth within the LSF's smoothing
    assert np.max(np.abs(sp.flux_model["total"] - clean)) < 0.1
    cen_true = physics.Freq2wave(truth[:, 1])
    sig_true = truth[:, 3] / 2.3548200450309493
    for k in range(2):
        j = int(np.argmin(np.abs(params["centers"] - cen_true[k])))
        assert abs(params["centers"][j] - cen_true[k]) < 0.5 * (wave[1] - wave[0])
        assert params["b"][j] == pytest.approx(physics.DopplerParameter(sig_true[k], 1548.204), rel=0.15)
        # central optical depth of the first line (amplitude_L alone is degenerate with the poorly constrained L / G)
        th = sp.regions[k].theta[0]
        assert th[0] * synth.voigt_peak_shape(th[2], th[3]) == pytest.approx(
            truth[k, 0] * synth.voigt_peak_shape(truth[k, 2], truth[k, 3]), rel=0.15)
        assert params["N"][j] > 0 and params["N_std"][j] > 0 and params["EW"][j] > 0
