"""Synthetic absorption regions for the BASELINE.json configs.

Pure numpy, float64, counter-based ``numpy.random.Philox`` keyed by the *global
region id* so that every rank (and the CPU oracle) generates identical inputs.
The flux recipe follows ``vamp_1.0/vpfits_intro.ipynb`` cell 7:
``flux = Tau2flux(sum tau) + N(0, sigma)``, negatives clipped to 0, constant
noise vector.  Profiles are evaluated with ``scipy.special.wofz`` - this is
workload generation, not the product path.
"""
import numpy as np
from scipy.special import wofz

SEED_BASE = 0x56414D50  # "VAMP"
C_REF = 2.98e8          # vamp_1.0/physics.py:3 (the reference's own value)
DNU = 4.0213e10         # Hz: 5 km/s at 1215.67 A with C_REF (matches simba_H1215.h5)
NU_LYA = C_REF / (1215.67e-10)
FWHM_PER_SIGMA = 2.0 * np.sqrt(2.0 * np.log(2.0))


def rng_for(region_id, stream=0):
    return np.random.Generator(np.random.Philox(key=[(SEED_BASE ^ int(region_id)) & 0xFFFFFFFFFFFFFFFF, int(stream)]))


def voigt_tau(x, amplitude, centroid, L_fwhm, G_fwhm):
    z = (2.0 * (x - centroid) + 1j * L_fwhm) * np.sqrt(np.log(2.0)) / G_fwhm
    return wofz(z).real * np.sqrt(np.pi * np.log(2.0)) / G_fwhm * L_fwhm * amplitude


def gauss_tau(x, amplitude, centroid, sigma):
    return amplitude * np.exp(-0.5 * ((x - centroid) / sigma) ** 2)


def frequency_grid(P):
    """Ascending frequency grid centred on Ly-alpha, DNU per pixel."""
    return NU_LYA + (np.arange(P, dtype=np.float64) - (P - 1) / 2.0) * DNU


def voigt_peak_shape(L_fwhm, G_fwhm):
    """H(0, y) = sqrt(pi) y erfcx(y): tau at line centre per unit ``amplitude_L``."""
    y = np.sqrt(np.log(2.0)) * L_fwhm / G_fwhm
    return np.sqrt(np.pi) * y * wofz(1j * y).real


def make_region(P, n_true, snr, region_id, voigt=True, lsf=None, ratio_range=(0.01, 0.3)):
    """One synthetic region.  Returns a dict with ``nu`` (Hz, ascending), ``flux``,
    ``noise`` (constant ``1/snr`` vector), ``truth`` ``[n_true, 4]`` =
    ``(amplitude, centroid, L_fwhm | sigma, G_fwhm | 0)`` in Hz and ``clean``.

    The central optical depth tau_0 ~ U(0.2, 3); for Voigt components the
    model's ``amplitude_L`` is tau_0 / H(0, y) (``tau0`` holds the drawn central depths).
    ``ratio_range``: log-uniform range of L_fwhm / G_fwhm."""
    rng = rng_for(region_id)
    nu = frequency_grid(P)
    sig_px = rng.uniform(3.0, 10.0, n_true)
    if P < 64:
        sig_px = rng.uniform(1.0, max(1.5, P / 12.0), n_true)
    amp = rng.uniform(0.2, 3.0, n_true)
    lo, hi = 0.1 * (P - 1), 0.9 * (P - 1)
    cen_px = np.empty(n_true)
    for k in range(n_true):
        for _ in range(200):
            c = rng.uniform(lo, hi)
            if all(abs(c - cen_px[j]) >= 1.5 * max(sig_px[k], sig_px[j]) for j in range(k)):
                break
        cen_px[k] = c
    ratio = 10.0 ** rng.uniform(np.log10(ratio_range[0]), np.log10(ratio_range[1]), n_true)
    truth = np.zeros((n_true, 4))
    truth[:, 0] = amp
    truth[:, 1] = nu[0] + cen_px * DNU
    tau = np.zeros(P)
    for k in range(n_true):
        if voigt:
            G = FWHM_PER_SIGMA * sig_px[k] * DNU
            truth[k, 2], truth[k, 3] = ratio[k] * G, G
            truth[k, 0] = amp[k] / voigt_peak_shape(truth[k, 2], truth[k, 3])
            tau += voigt_tau(nu, truth[k, 0], truth[k, 1], truth[k, 2], truth[k, 3])
        else:
            truth[k, 2] = sig_px[k] * DNU
            tau += gauss_tau(nu, amp[k], truth[k, 1], truth[k, 2])
    clean = np.exp(-tau)
    if lsf is not None:
        lsf = np.asarray(lsf, dtype=np.float64)
        h = lsf.size // 2
        clean = 1.0 + np.convolve(np.concatenate([np.zeros(h), clean - 1.0, np.zeros(h)]), lsf, "valid")
    sigma = 1.0 / float(snr)
    flux = clean + rng.normal(0.0, sigma, P)
    flux[flux < 0.0] = 0.0
    return {"nu": nu, "flux": flux, "noise": np.full(P, sigma), "truth": truth, "clean": clean,
            "voigt": bool(voigt), "region_id": int(region_id), "snr": float(snr), "tau0": amp}


def make_doublet_region(P, n_true, snr, region_id, doublet="CIV", voigt=True, lsf=None, dv_kms=5.0):
    """One synthetic region holding BOTH lines of a doublet (BASELINE config C3: CIV 1548/1550, MgII 2796/2803;
    tied velocity, Doppler width and column density): ``n_true`` absorbers, each the sum of its line and the tied
    second line (``vamp_b200.physics.doublet_ratios``).  The grid has ``dv_kms`` pixels around the doublet; ``truth
    [n_true, 4]`` holds the FIRST line's ``(amplitude, centroid, L_fwhm | sigma, G_fwhm | 0)`` in Hz, ``multiplet``
    the ratios to pass to ``Regions(multiplet=...)`` / ``fit_regions(multiplet=...)``."""
    from . import physics
    fr, ar, lr = physics.doublet_ratios(doublet)
    rng = rng_for(region_id, stream=2)
    nu1 = C_REF / (physics.DOUBLETS[doublet][0][0] * 1e-10)
    dnu = nu1 * dv_kms * 1e3 / C_REF
    sep = nu1 * (1.0 - fr) / dnu                                # separation of the two lines in pixels
    if sep + 60 > P - 1:
        raise ValueError("region of %d px cannot hold a doublet %.0f px apart" % (P, sep))
    p2 = 0.5 * (P - 1 - sep)                                    # rest position of line 2; line 1 at p2 + sep
    nu = nu1 + (np.arange(P, dtype=np.float64) - (p2 + sep)) * dnu
    half = min(p2 - 12.0, 0.45 * sep)                           # absorbers within +- half px of the rest position
    sig_px = rng.uniform(2.5, 6.0, n_true)
    tau0 = rng.uniform(0.4, 2.5, n_true)
    off = np.empty(n_true)
    for k in range(n_true):
        for _ in range(200):
            o = rng.uniform(-half, half)
            if all(abs(o - off[j]) >= 2.0 * max(sig_px[k], sig_px[j]) for j in range(k)):
                break
        off[k] = o
    ratio = 10.0 ** rng.uniform(-2.0, np.log10(0.3), n_true)
    truth = np.zeros((n_true, 4))
    truth[:, 1] = nu1 + off * dnu
    tau = np.zeros(P)
    for k in range(n_true):
        c = truth[k, 1]
        if voigt:
            G = FWHM_PER_SIGMA * sig_px[k] * dnu
            truth[k, 2], truth[k, 3] = ratio[k] * G, G
            truth[k, 0] = tau0[k] / voigt_peak_shape(truth[k, 2], truth[k, 3])
            tau += voigt_tau(nu, truth[k, 0], c, truth[k, 2], G)
            tau += voigt_tau(nu, ar * truth[k, 0], fr * c, lr * truth[k, 2], fr * G)
        else:
            truth[k, 0], truth[k, 2] = tau0[k], sig_px[k] * dnu
            tau += gauss_tau(nu, tau0[k], c, truth[k, 2])
            tau += gauss_tau(nu, ar * tau0[k], fr * c, fr * truth[k, 2])
    clean = np.exp(-tau)
    if lsf is not None:
        lsf = np.asarray(lsf, dtype=np.float64)
        h = lsf.size // 2
        clean = 1.0 + np.convolve(np.concatenate([np.zeros(h), clean - 1.0, np.zeros(h)]), lsf, "valid")
    sigma = 1.0 / float(snr)
    flux = clean + rng.normal(0.0, sigma, P)
    flux[flux < 0.0] = 0.0
    return {"nu": nu, "flux": flux, "noise": np.full(P, sigma), "truth": truth, "clean": clean, "voigt": bool(voigt),
            "region_id": int(region_id), "snr": float(snr), "tau0": tau0, "multiplet": (fr, ar, lr), "dnu": dnu,
            "sep_px": sep}


def make_sightline(ion="HI", P=1000, snr=70.0, systems=((200.0, 1.6, 5.0), (450.0, 0.9, 6.5), (700.0, 1.3, 4.0)), lsf=None,
                   seed=5, dv_kms=5.0, lorentz_over_gauss=0.05):
    """One SIMBA-like sightline of ``P`` velocity pixels around one ion's line (BASELINE config C3; the reference's
    fixtures ``simba_H1215.h5`` / ``simba_CII1036.h5`` are such per-ion 1000-pixel spectra at 5 km/s).  ``ion``: "HI"
    (Ly-alpha 1215.67, single line) or a doublet of ``physics.DOUBLETS`` ("CIV", "MgII": both lines of every absorber,
    tied velocity / Doppler width / column density).  ``systems``: ``(pixel of the first line, tau_0, sigma in
    pixels)`` per absorber.  Returns a dict: ``wave`` (Angstrom, ascending), ``flux``, ``noise``, ``clean`` (noiseless,
    LSF-convolved), ``clean_intrinsic``, ``truth [n, 4]`` = the first line's (amplitude_L, centroid, L_fwhm, G_fwhm) in Hz,
    ``line``, ``doublet`` (None for HI), ``multiplet``."""
    from . import physics
    rng = np.random.Generator(np.random.Philox(key=[SEED_BASE ^ 0xC3, int(seed)]))
    if ion == "HI":
        l1, doublet, mult = 1215.67, None, None
    else:
        l1, doublet, mult = physics.DOUBLETS[ion][0][0], ion, physics.doublet_ratios(ion)
    wave = l1 * (1.0 - 0.004) + (l1 * dv_kms * 1.0e3 / C_REF) * np.arange(P)
    nu = C_REF / (wave * 1e-10)
    truth, tau = [], np.zeros(P)
    for px, tau0, sig_px in systems:
        i = int(px)
        c = nu[i]
        G = FWHM_PER_SIGMA * sig_px * abs(nu[i] - nu[i + 1])
        L = lorentz_over_gauss * G
        A = tau0 / voigt_peak_shape(L, G)
        tau += voigt_tau(nu, A, c, L, G)
        if mult is not None:
            fr, ar, lr = mult
            tau += voigt_tau(nu, ar * A, fr * c, lr * L, fr * G)
        truth.append((A, c, L, G))
    intrinsic = np.exp(-tau)
    clean = intrinsic
    if lsf is not None:
        lsf = np.asarray(lsf, dtype=np.float64)
        h = lsf.size // 2
        clean = 1.0 + np.convolve(np.concatenate([np.zeros(h), intrinsic - 1.0, np.zeros(h)]), lsf, "valid")
    sigma = 1.0 / float(snr)
    flux = clean + rng.normal(0.0, sigma, P)
    return {"ion": ion, "line": l1, "doublet": doublet, "multiplet": mult, "wave": wave, "flux": flux,
            "noise": np.full(P, sigma), "clean": clean, "clean_intrinsic": intrinsic, "truth": np.array(truth),
            "systems": [tuple(s_) for s_ in systems], "dv_kms": dv_kms, "snr": float(snr)}


def config_c3(lsf=True):
    """C3: one synthetic SIMBA-like sightline seen in three ions - HI 1215 (single line), CIV 1548/1550 and MgII
    2796/2803 (doublets with tied v / b / N) - each a 1000-pixel spectrum at 5 km/s, convolved with the Gaussian LSF
    (FWHM 2 px, 7 taps), S/N 70.  The same three absorbers (same velocities) in every ion, with
    ion-dependent central depths."""
    taps = gaussian_lsf(2.0, 7) if lsf else None
    base = ((200.0, 5.0), (450.0, 6.5), (700.0, 4.0))
    depth = {"HI": (2.2, 1.4, 1.8), "CIV": (1.6, 0.9, 1.3), "MgII": (1.2, 1.5, 0.8)}
    out = []
    for k, ion in enumerate(("HI", "CIV", "MgII")):
        systems = tuple((px, depth[ion][i], sig) for i, (px, sig) in enumerate(base))
        out.append(make_sightline(ion, 1000, 70.0, systems, lsf=taps, seed=30 + k))
        out[-1]["lsf"] = taps
    return out


def gaussian_lsf(fwhm_px=2.0, taps=7):
    """Normalised Gaussian LSF (config C3: FWHM 2 px, 7 taps)."""
    s = fwhm_px / FWHM_PER_SIGMA
    t = np.arange(taps) - taps // 2
    w = np.exp(-0.5 * (t / s) ** 2)
    return w / w.sum()


def config_c1():
    """C1: one HI Ly-alpha region, 1 Voigt component, 200 px, SNR 50, seed 1."""
    return make_region(200, 1, 50.0, region_id=1, voigt=True)


def config_c2(region_id=2, n_true=3):
    """C2: blended 256-px region (BIC sweep over n = 1..6 with 4096 chains each)."""
    return make_region(256, n_true, 50.0, region_id=region_id, voigt=True)


def config_c4_region(region_id):
    """C4: one of the 10^5 mixed regions: P ~ U{32..256}, n ~ U{1..8}, SNR in {10,30,100}."""
    r = rng_for(region_id, stream=1)
    P = int(r.integers(32, 257))
    n_true = int(r.integers(1, 9))
    n_true = max(1, min(n_true, P // 24))
    snr = float(r.choice([10.0, 30.0, 100.0]))
    return make_region(P, n_true, snr, region_id=region_id, voigt=True)


def prior_draw_theta(nu, n, voigt, rng, posterior_like=None):
    """Parameter vectors for throughput runs: prior draws
    (A ~ Gamma(2,1), c ~ U, widths ~ U(0, w_max)) or, with ``posterior_like`` =
    a truth array, truth +- 5 %.  Returns ``[n, 4]`` in Hz."""
    span = nu[-1] - nu[0]
    sigma_max = span / 2.0
    w_max = sigma_max * FWHM_PER_SIGMA if voigt else sigma_max
    th = np.zeros((n, 4))
    if posterior_like is not None:
        t = np.asarray(posterior_like)
        idx = rng.integers(0, t.shape[0], n)
        th[:] = t[idx] * (1.0 + rng.uniform(-0.05, 0.05, (n, 4)))
        th[:, 1] = t[idx, 1] + rng.uniform(-0.05, 0.05, n) * t[idx, 3 if voigt else 2]
        return th
    th[:, 0] = rng.gamma(2.0, 1.0, n)
    th[:, 1] = rng.uniform(nu[0], nu[-1], n)
    th[:, 2] = rng.uniform(0.0, w_max, n)
    if voigt:
        th[:, 3] = rng.uniform(0.0, w_max, n)
    return th
