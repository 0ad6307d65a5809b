"""GPU: K1 (fused tau -> exp(-sum tau) -> LSF -> chi^2 / lnL) against the oracle and the golden
vectors produced by the reference's own code, through the C ABI (device- and host-buffer entry
points).  Tolerances are north_star's: flux 1e-6 relative (fp64 path) / 1e-4 (fp32 path),
lnL 1e-6 relative (fp64 path)."""
import os

import numpy as np
import pytest
from scipy.special import wofz

from oracle import density as D
from vamp_b200 import engine, synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FLUX_RTOL = {"f64": 1e-6, "f32": 1e-4}
FLUX_ATOL = {"f64": 1e-290, "f32": 1e-30}      # exp(-tau) underflows to denormals/0 for tau > ~87 in fp32
LNL_RTOL = {"f64": 1e-6, "f32": 2e-4}          # fp32 lnL is reported, not part of the 1e-6 claim
CHI_RTOL = {"f64": 1e-6, "f32": 2e-3}


def test_philox_known_answers(ctx):
    # Random123 kat_vectors, philox4x32-10
    assert engine.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert engine.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert engine.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("dtype,tol", [("f32", 1.5e-6), ("f64", 5e-14)])
def test_faddeeva_plane(ctx, dtype, tol):
    rng = np.random.default_rng(0)
    n = 200000
    x = np.concatenate([rng.uniform(0, 16, n), 10 ** rng.uniform(-3, 5, n), -rng.uniform(0, 12, n // 4)])
    y = np.concatenate([10 ** rng.uniform(-6, 1.5, n), 10 ** rng.uniform(-6, 5, n), 10 ** rng.uniform(-3, 1, n // 4)])
    H = engine.faddeeva_H(x, y, dtype)
    ref = np.sqrt(np.pi) * y * wofz(x + 1j * y).real
    err = np.abs(H - ref)
    assert np.all(np.isfinite(H))
    assert err.max() < tol, "max |H - H_wofz| = %.3e at x=%g y=%g" % (err.max(), x[err.argmax()], y[err.argmax()])


@pytest.mark.parametrize("dtype,tol", [("f32", 1.5e-6), ("f64", 5e-14)])
def test_faddeeva_mpmath_golden(ctx, dtype, tol):
    g = np.load(os.path.join(GOLDEN, "wofz_mp.npz"))
    H = engine.faddeeva_H(g["x"], g["y"], dtype)
    ref = np.sqrt(np.pi) * g["y"] * g["re"]
    assert np.abs(H - ref).max() < tol


def test_faddeeva_extremes(ctx):
    # G -> 0 (x, y -> huge), L -> 0 (y -> 0), exact zeros: finite, correct limits
    x = np.array([0.0, 0.0, 1e15, 1e10, 3.0, 0.0, 1e-30])
    y = np.array([1e30, 0.0, 1e15, 1e-10, 0.0, 1e-30, 1e-30])
    for dtype in ("f32", "f64"):
        H = engine.faddeeva_H(x, y, dtype)
        assert np.all(np.isfinite(H)), (dtype, H)
        assert H[0] == pytest.approx(1.0, abs=1e-6) and H[1] == 0.0 and H[2] == pytest.approx(0.5, abs=1e-6)
        assert abs(H[3]) < 1e-30 and H[4] == 0.0


def _run_case(ctx, g, name, dtype, host):
    nu, flux, noise = g[name + "/nu"], g[name + "/flux"], g[name + "/noise"]
    comps, voigt, sd = g[name + "/comps"], bool(g[name + "/voigt"]), float(g[name + "/sd"])
    regs = engine.Regions(ctx, [(nu, flux, noise)])
    fn = engine.loglike_host if host else engine.loglike
    lnL_sd, chi2, mflux = fn(regs, comps[None], [0], sd=[sd], voigt=voigt, dtype=dtype, want_flux=True)
    lnL_nm, _ = fn(regs, comps[None], [0], sd=None, voigt=voigt, dtype=dtype)
    return lnL_sd[0], lnL_nm[0], chi2[0], mflux[0, : len(nu)]


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("host", [False, True])
def test_golden_cases(ctx, golden_profiles, dtype, host):
    g = golden_profiles
    worst = 0.0
    for name in [str(n) for n in g["names"]]:
        lnL_sd, lnL_nm, chi2, mflux = _run_case(ctx, g, name, dtype, host)
        total = g[name + "/total"]
        np.testing.assert_allclose(mflux, total, rtol=FLUX_RTOL[dtype], atol=FLUX_ATOL[dtype], err_msg=name)
        worst = max(worst, np.max(np.abs(mflux - total) / np.maximum(total, 1e-30)))
        ref_chi = float(g[name + "/chi2"])
        assert chi2 == pytest.approx(ref_chi, rel=CHI_RTOL[dtype], abs=1e-9), name
        ref_nm = float(g[name + "/lnL_noise_map"])
        assert lnL_nm == pytest.approx(ref_nm, rel=LNL_RTOL[dtype]), name
        ref_sd = D.normal_loglike(g[name + "/flux"], total, sd=float(g[name + "/sd"]))
        assert lnL_sd == pytest.approx(ref_sd, rel=LNL_RTOL[dtype]), name
    print("worst relative flux error (%s): %.3e" % (dtype, worst))


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("voigt", [True, False])
def test_batched_ragged_regions(ctx, dtype, voigt):
    """Many regions of different sizes (1..3000 px), many items per region, ragged ncomp."""
    rng = np.random.default_rng(11)
    sizes = [1, 2, 5, 17, 32, 33, 64, 100, 128, 200, 256, 257, 600, 2048, 2049, 3000]
    regs_h = [synth.make_region(P, 1 + P % 3, 30.0, 500 + P, voigt=voigt) if P >= 8 else
              {"nu": synth.frequency_grid(8)[:P], "flux": rng.uniform(0.2, 1, P), "noise": np.full(P, 0.05)} for P in sizes]
    regs = engine.Regions(ctx, [(r["nu"], r["flux"], r["noise"]) for r in regs_h])
    cmax, B = 7, 96
    item_region = rng.integers(0, len(sizes), B)
    item_ncomp = rng.integers(0, cmax + 1, B)
    comps = np.zeros((B, cmax, 4))
    sd = rng.uniform(0.01, 0.5, B)
    for b in range(B):
        nu = regs_h[item_region[b]]["nu"]
        if len(nu) >= 3:
            comps[b, : item_ncomp[b]] = synth.prior_draw_theta(nu, cmax, voigt, rng)[: item_ncomp[b]]
        else:
            comps[b, : item_ncomp[b]] = [1.0, nu[0], 3 * synth.DNU, 5 * synth.DNU]
    lnL, chi2, mflux = engine.loglike(regs, comps, item_region, item_ncomp, sd=sd, voigt=voigt, dtype=dtype, want_flux=True)
    lnL_h, chi2_h = engine.loglike_host(regs, comps, item_region, item_ncomp, sd=sd, voigt=voigt, dtype=dtype)
    np.testing.assert_array_equal(lnL, lnL_h)      # the two ABI entry points run the same kernels
    np.testing.assert_array_equal(chi2, chi2_h)
    for b in range(B):
        r = regs_h[item_region[b]]
        ref = D.model_flux(r["nu"], comps[b, : item_ncomp[b]], voigt) if item_ncomp[b] else np.ones(len(r["nu"]))
        np.testing.assert_allclose(mflux[b, : len(ref)], ref, rtol=FLUX_RTOL[dtype], atol=FLUX_ATOL[dtype], err_msg=str(b))
        assert np.all(mflux[b, len(ref):] == 0.0)
        assert chi2[b] == pytest.approx(D.chisquared(r["flux"], ref, r["noise"]), rel=CHI_RTOL[dtype], abs=1e-9)
        assert lnL[b] == pytest.approx(D.normal_loglike(r["flux"], ref, sd=sd[b]), rel=LNL_RTOL[dtype])


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("taps", [3, 7, 33])
def test_lsf_against_oracle(ctx, dtype, taps):
    rng = np.random.default_rng(taps)
    lsf = synth.gaussian_lsf(fwhm_px=taps / 3.0, taps=taps)
    if taps == 3:
        lsf = np.array([0.5, 0.3, 0.2])          # asymmetric: checks convolution orientation
    sizes = [5, 40, 256, 257, 2047, 2048, 2050, 5000]
    regs_h = [synth.make_region(P, 2, 50.0, 700 + P, voigt=True, lsf=lsf) if P >= 8 else
              {"nu": synth.frequency_grid(8)[:P], "flux": rng.uniform(0.2, 1, P), "noise": np.full(P, 0.05)} for P in sizes]
    regs = engine.Regions(ctx, [(r["nu"], r["flux"], r["noise"]) for r in regs_h], lsf=lsf)
    comps = np.zeros((len(sizes), 3, 4))
    for i, r in enumerate(regs_h):
        comps[i] = synth.prior_draw_theta(r["nu"], 3, True, rng)
        comps[i, :, 2:] = np.minimum(comps[i, :, 2:], 30 * synth.DNU) + synth.DNU   # keep features narrow: LSF matters
    lnL, chi2, mflux = engine.loglike(regs, comps, np.arange(len(sizes)), sd=None, voigt=True, dtype=dtype, want_flux=True)
    for i, r in enumerate(regs_h):
        ref = D.model_flux(r["nu"], comps[i], True, lsf)
        np.testing.assert_allclose(mflux[i, : len(ref)], ref, rtol=FLUX_RTOL[dtype] * 3, atol=1e-12, err_msg=str(sizes[i]))
        assert lnL[i] == pytest.approx(D.noise_map_loglike(r["flux"], r["noise"], ref), rel=LNL_RTOL[dtype])


def test_identity_lsf_is_bit_identical(ctx):
    r = synth.config_c2()
    comps = r["truth"][None]
    a = engine.loglike(engine.Regions(ctx, [(r["nu"], r["flux"], r["noise"])]), comps, [0], sd=[0.02], dtype="f64", want_flux=True)
    b = engine.loglike(engine.Regions(ctx, [(r["nu"], r["flux"], r["noise"])], lsf=[1.0]), comps, [0], sd=[0.02], dtype="f64", want_flux=True)
    for u, v in zip(a, b):
        np.testing.assert_array_equal(u, v)


def test_full_size_properties(ctx):
    """BASELINE config-2 size (4096 chains x n=1..6 on a 256-px region): size-independent
    properties - permutation invariance, additivity of tau, zero-amplitude components."""
    r = synth.config_c2()
    regs = engine.Regions(ctx, [(r["nu"], r["flux"], r["noise"])])
    rng = np.random.default_rng(2)
    B, cmax = 4096 * 6, 6
    ncomp = np.repeat(np.arange(1, 7), 4096).astype(np.int32)
    comps = np.zeros((B, cmax, 4))
    pri = synth.prior_draw_theta(r["nu"], B * cmax, True, rng).reshape(B, cmax, 4)
    for n in range(1, 7):
        comps[ncomp == n, :n] = pri[ncomp == n, :n]
    sd = rng.uniform(0.01, 0.3, B)
    lnL, chi2 = engine.loglike(regs, comps, np.zeros(B, np.int32), ncomp, sd=sd, dtype="f64")
    assert np.all(np.isfinite(lnL)) and np.all(chi2 >= 0)
    # (a) component order does not matter
    perm = comps.copy()
    for n in range(2, 7):
        perm[ncomp == n, :n] = perm[ncomp == n, :n][:, ::-1]
    lnL_p, chi2_p = engine.loglike(regs, perm, np.zeros(B, np.int32), ncomp, sd=sd, dtype="f64")
    np.testing.assert_allclose(lnL_p, lnL, rtol=1e-11)
    # (b) an extra zero-amplitude component changes nothing (padding slot used as a real slot)
    ext = np.zeros((B, cmax + 1, 4))
    ext[:, :cmax] = comps
    ncomp_e = ncomp + 1
    for n in range(1, 7):
        ext[ncomp == n, n] = [0.0, r["nu"][100], 3 * synth.DNU, 9 * synth.DNU]
    lnL_e, _ = engine.loglike(regs, ext, np.zeros(B, np.int32), ncomp_e, sd=sd, dtype="f64")
    np.testing.assert_allclose(lnL_e, lnL, rtol=1e-12)
    # (c) spot-check 64 random items against the oracle
    for b in rng.integers(0, B, 64):
        ref = D.normal_loglike(r["flux"], D.model_flux(r["nu"], comps[b, : ncomp[b]], True), sd=sd[b])
        assert lnL[b] == pytest.approx(ref, rel=1e-6)
    # (d) fp32 path agrees with fp64 on the whole batch
    lnL32, chi32 = engine.loglike(regs, comps, np.zeros(B, np.int32), ncomp, sd=sd, dtype="f32")
    np.testing.assert_allclose(chi32, chi2, rtol=2e-3, atol=1e-3)


def test_profile_tau_matches_reference_voigt(ctx, golden_profiles):
    g = golden_profiles
    for name in ("c1_truth_voigt", "edge_saturated", "edge_tiny_L", "edge_tiny_G", "edge_g_narrow"):
        nu, comps, voigt = g[name + "/nu"], g[name + "/comps"], bool(g[name + "/voigt"])
        for k, (a, c, w1, w2) in enumerate(comps):
            tau = engine.profile_tau(nu, c, a, w1, w2, voigt=voigt, ctx=ctx)
            np.testing.assert_allclose(tau, g[name + "/taus"][k], rtol=1e-9, atol=1e-13 * max(a, 1.0), err_msg=name)


def test_c5_largest_corner(ctx):
    """BASELINE config C5's largest shapes - 65 536-pixel regions x 32 components - through the chunked CTA schedule:
    oracle comparison on a few items, then size-independent properties on a batch (multiplicativity of the flux in
    disjoint component sets, permutation invariance, lnL from the returned flux)."""
    P, C = 65536, 32
    rng = np.random.default_rng(65536)
    nu = synth.NU_LYA + (np.arange(P) - P / 2.0) * synth.DNU
    cen = rng.uniform(nu[200], nu[-200], C)
    G = rng.uniform(4.0, 40.0, C) * synth.DNU
    comps = np.stack([rng.uniform(0.2, 5.0, C), cen, G * 10 ** rng.uniform(-2, -0.3, C), G], axis=1)
    clean = D.model_flux(nu, comps, True)
    flux = clean + rng.normal(0, 0.02, P)
    regs = engine.Regions(ctx, [(nu, flux, np.full(P, 0.02))])
    for dtype in ("f64", "f32"):
        lnL, chi2, m = engine.loglike(regs, comps[None], [0], sd=[0.02], dtype=dtype, want_flux=True)
        np.testing.assert_allclose(m[0], clean, rtol=FLUX_RTOL[dtype], atol=FLUX_ATOL[dtype])
        if dtype == "f64":
            assert lnL[0] == pytest.approx(D.normal_loglike(flux, clean, sd=0.02), rel=LNL_RTOL[dtype])
            assert chi2[0] == pytest.approx(D.chisquared(flux, clean, np.full(P, 0.02)), rel=CHI_RTOL[dtype])
    # a batch of 48 parameter vectors (fp64): flux(all 32) = flux(first 16) x flux(last 16); order does not matter
    B = 16
    th = np.repeat(comps[None], B, axis=0) * (1.0 + rng.uniform(-0.2, 0.2, (B, C, 4)))
    th[:, :, 1] = comps[None, :, 1] + rng.uniform(-20, 20, (B, C)) * synth.DNU
    halves = np.zeros((2 * B, C, 4))
    halves[:B, :16], halves[B:, :16] = th[:, :16], th[:, 16:]
    ir = np.zeros(B, np.int32)
    lnL_a, _, m_all = engine.loglike(regs, th, ir, sd=np.full(B, 0.05), dtype="f64", want_flux=True)
    _, _, m_half = engine.loglike(regs, halves, np.zeros(2 * B, np.int32), np.full(2 * B, 16, np.int32), sd=np.full(2 * B, 0.05),
                                  dtype="f64", want_flux=True)
    np.testing.assert_allclose(m_half[:B] * m_half[B:], m_all, rtol=1e-12, atol=1e-300)
    lnL_p, _ = engine.loglike(regs, th[:, ::-1].copy(), ir, sd=np.full(B, 0.05), dtype="f64")
    np.testing.assert_allclose(lnL_p, lnL_a, rtol=1e-11)
    want = [D.normal_loglike(flux, m_all[b], sd=0.05) for b in range(B)]
    np.testing.assert_allclose(lnL_a, want, rtol=1e-9)
