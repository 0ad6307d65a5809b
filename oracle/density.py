"""Oracle: the target density of the reference fit, restated in numpy float64.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites the
reference ``file:line`` (paths relative to ``/root/reference``) it restates.

Parameter convention used throughout the oracle and the tests
--------------------------------------------------------------
``comps`` is a float64 array ``[n, 4]`` in the reference's physical units (Hz):

* Gaussian mode: ``(amplitude, centroid, sigma, unused)``   - ``vpfits.py:239-252``
* Voigt mode:    ``(amplitude, centroid, L_fwhm, G_fwhm)``  - ``vpfits.py:283-297``

``sd`` is the free noise parameter of ``VPfit(noise=None)`` (``vpfits.py:39``).
"""
import numpy as np
from scipy.special import wofz, voigt_profile

SQRT_LN2 = np.sqrt(np.log(2.0))
SQRT_PI_LN2 = np.sqrt(np.pi * np.log(2.0))
FWHM_PER_SIGMA = 2.0 * np.sqrt(2.0 * np.log(2.0))


# --------------------------------------------------------------------------- profiles
def gauss_tau(x, amplitude, centroid, sigma):
    """``VPfit.GaussFunction`` - vamp_1.0/vpfits.py:44-54."""
    x = np.asarray(x, dtype=np.float64)
    return amplitude * np.exp(-0.5 * ((x - centroid) / sigma) ** 2)


def voigt_tau(x, centroid, amplitude, L_fwhm, G_fwhm):
    """``VPfit.VoigtFunction`` - vamp_1.0/vpfits.py:58-76.

    The reference builds ``astropy.modeling.models.Voigt1D(x_0, amplitude_L, fwhm_L,
    fwhm_G)`` and calls it.  astropy is a third-party dependency absent from
    /root/reference (unpinned in vamp_1.0/requirements.txt); north_star fixes the
    oracle to its ``scipy.special.wofz`` evaluation:

        z   = (2 (x - x_0) + i fwhm_L) sqrt(ln 2) / fwhm_G
        tau = Re w(z) * sqrt(pi ln 2) * fwhm_L * amplitude_L / fwhm_G
    """
    x = np.asarray(x, dtype=np.float64)
    z = (2.0 * (x - centroid) + 1j * L_fwhm) * SQRT_LN2 / G_fwhm
    return wofz(z).real * SQRT_PI_LN2 / G_fwhm * L_fwhm * amplitude


def voigt_tau_via_voigt_profile(x, centroid, amplitude, L_fwhm, G_fwhm):
    """Independent second evaluation of the same profile:
    ``A (pi L / 2) voigt_profile(x - c, G / (2 sqrt(2 ln 2)), L / 2)``."""
    x = np.asarray(x, dtype=np.float64)
    return amplitude * (np.pi * L_fwhm / 2.0) * voigt_profile(
        x - centroid, G_fwhm / FWHM_PER_SIGMA, L_fwhm / 2.0)


def gaussian_width(G_fwhm):
    """``VPfit.GaussianWidth`` - vamp_1.0/vpfits.py:80-88."""
    return G_fwhm / (2.0 * np.sqrt(2.0 * np.log(2.0)))


def tau2flux(tau):
    """``physics.Tau2flux`` - vamp_1.0/physics.py:98-105."""
    return np.exp(-tau)


def second_line(comp, voigt, multiplet):
    """Tied second line of a component (doublets of BASELINE config C3; NOT IN THE REFERENCE, which
    evaluates ``VoigtFunction`` once per component - parity-unpinned convention of this repo, stated in
    include/vamp_b200.h ``vamp_regions_set_multiplet``): with ``multiplet = (freq_ratio, amp_ratio,
    lorentz_ratio[, shift])`` the line has amplitude ``amp_ratio A``, centroid ``freq_ratio c + shift``
    (``shift`` = 0 in Hz; non-zero only in the kernels' scaled coordinates), Gaussian width
    ``freq_ratio G`` (sigma likewise) and Lorentzian width ``lorentz_ratio L``."""
    fr, ar, lr = multiplet[0], multiplet[1], multiplet[2]
    sh = multiplet[3] if len(multiplet) > 3 else 0.0
    a, c, w1, w2 = comp
    return (ar * a, fr * c + sh, lr * w1, fr * w2) if voigt else (ar * a, fr * c + sh, fr * w1, w2)


def component_taus(x, comps, voigt, multiplet=None):
    """Per-component optical depths = the ``component_i`` deterministics
    (vamp_1.0/vpfits.py:254-262, 299-307).  Returns ``[n, P]``.  ``multiplet``: every component is the
    sum of its line and the tied :func:`second_line`, each evaluated by the reference's profile function."""
    comps = np.atleast_2d(np.asarray(comps, dtype=np.float64))
    out = np.empty((comps.shape[0], len(x)), dtype=np.float64)
    for k, (a, c, w1, w2) in enumerate(comps):
        out[k] = voigt_tau(x, c, a, w1, w2) if voigt else gauss_tau(x, a, c, w1)
        if multiplet is not None and multiplet[1] != 0.0:
            a2, c2, v1, v2 = second_line((a, c, w1, w2), voigt, multiplet)
            with np.errstate(all="ignore"):
                t2 = voigt_tau(x, c2, a2, v1, v2) if voigt else gauss_tau(x, a2, c2, v1)
            out[k] += t2
    return out


def lsf_convolve(flux, lsf):
    """Instrumental line-spread function applied to ``exp(-tau)``.

    NOT IN THE REFERENCE ("absent in both generations"); the
    convention is this repo's and is parity-unpinned: a true convolution of the
    flux *decrement* with zero padding, i.e. the continuum (flux 1) extends
    beyond the region:  ``1 + np.convolve(flux - 1, lsf, 'same')``.
    ``lsf=None`` or ``[1.0]`` is the identity and reproduces the reference."""
    if lsf is None:
        return np.asarray(flux, dtype=np.float64)
    lsf = np.asarray(lsf, dtype=np.float64)
    if lsf.size % 2 != 1:
        raise ValueError("LSF must have an odd number of taps")
    flux = np.asarray(flux, dtype=np.float64)
    if lsf.size == 1 and lsf[0] == 1.0:
        return flux
    h = lsf.size // 2
    padded = np.concatenate([np.zeros(h), flux - 1.0, np.zeros(h)])
    return 1.0 + np.convolve(padded, lsf, "valid")


def model_flux(x, comps, voigt, lsf=None, multiplet=None):
    """The ``total`` deterministic: ``Tau2flux(sum(profile_sum))`` -
    vamp_1.0/vpfits.py:334-338 (+ this repo's LSF and multiplets)."""
    taus = component_taus(x, comps, voigt, multiplet)
    return lsf_convolve(tau2flux(taus.sum(axis=0)), lsf)


# --------------------------------------------------------------------------- likelihood
def normal_loglike(flux, model, sd=None, noise=None):
    """``mc.Normal("obs", total, std_deviation, value=flux, observed=True)`` -
    vamp_1.0/vpfits.py:341 with the precision of vpfits.py:39,
    ``tau_p = 1/sd**2`` (sd free) or ``1/noise**2`` (``VPfit(noise=...)``).

    PyMC 2.3 ``normal_like(x, mu, tau) = sum(-tau/2 (x-mu)^2 + 1/2 ln(tau / 2 pi))``
    [third party, restated]."""
    flux = np.asarray(flux, dtype=np.float64)
    model = np.asarray(model, dtype=np.float64)
    if (sd is None) == (noise is None):
        raise ValueError("give exactly one of sd / noise")
    taup = 1.0 / np.square(sd if noise is None else np.asarray(noise, dtype=np.float64))
    taup = np.broadcast_to(taup, flux.shape)
    return float(np.sum(-0.5 * taup * (flux - model) ** 2 + 0.5 * np.log(taup / (2.0 * np.pi))))


def noise_map_loglike(data, noise_map, model):
    """vamp_2.0/vamp_src/fit/fit.py:83-171:
    ``-0.5 (sum(((d-m)/s)^2) + sum(ln(2 pi s^2)))``."""
    data = np.asarray(data, dtype=np.float64)
    noise_map = np.asarray(noise_map, dtype=np.float64)
    chi2 = np.sum(np.square((data - np.asarray(model, dtype=np.float64)) / noise_map))
    return float(-0.5 * (chi2 + np.sum(np.log(2.0 * np.pi * noise_map ** 2))))


def chisquared(observed, expected, noise):
    """``VPfit.Chisquared`` - vamp_1.0/vpfits.py:109-118."""
    return float(np.sum(((np.asarray(observed) - np.asarray(expected)) / np.asarray(noise)) ** 2))


def reduced_chisquared(observed, expected, noise, freedom):
    """``VPfit.ReducedChisquared`` - vamp_1.0/vpfits.py:121-131."""
    return chisquared(observed, expected, noise) / freedom


# --------------------------------------------------------------------------- priors
def prior_bounds(x, voigt):
    """Uniform boxes - vamp_1.0/vpfits.py:249-252 (Gauss), 292-297 (Voigt) with
    ``sigma_max`` / ``fwhm_max`` from vpfits.py:320,326.
    Returns ``(c_lo, c_hi, w_max)``: centroid in ``[x[0], x[-1]]``, widths in
    ``[0, w_max]`` (sigma_max for Gauss; fwhm_max for both Voigt widths)."""
    sigma_max = (x[-1] - x[0]) / 2.0
    w_max = sigma_max * 2.0 * np.sqrt(2.0 * np.log(2.0)) if voigt else sigma_max
    return float(x[0]), float(x[-1]), float(w_max)


def log_prior(comps, x, voigt, sd=None):
    """Sum of the stochastics' log-densities.

    * ``xexp`` amplitude: ``ln(v exp(-v))`` for v >= 0 else -inf -
      vamp_1.0/vpfits.py:239-244 / 283-288 (ln 0 = -inf at v = 0).
    * ``mc.Uniform`` centroid / sigma / L / G - vpfits.py:249-252, 292-297;
      PyMC 2 ``uniform_like = -ln(hi - lo)`` inside ``[lo, hi]``, -inf outside.
    * ``sd ~ Uniform(0, 1)`` - vpfits.py:39 (density 1)."""
    comps = np.atleast_2d(np.asarray(comps, dtype=np.float64))
    c_lo, c_hi, w_max = prior_bounds(x, voigt)
    lp = 0.0
    for a, c, w1, w2 in comps:
        if a < 0.0:
            return -np.inf
        with np.errstate(divide="ignore"):
            lp += np.log(a * np.exp(-a))
        if not (c_lo <= c <= c_hi):
            return -np.inf
        lp -= np.log(c_hi - c_lo)
        widths = (w1, w2) if voigt else (w1,)
        for w in widths:
            if not (0.0 <= w <= w_max):
                return -np.inf
            lp -= np.log(w_max)
    if sd is not None and not (0.0 <= sd <= 1.0):
        return -np.inf
    return float(lp)


def log_posterior(comps, x, flux, voigt, sd=None, noise=None, lsf=None, multiplet=None):
    """``model.logp`` = priors + observed Normal (what MCMC and MAP see)."""
    lp = log_prior(comps, x, voigt, sd=sd)
    if not np.isfinite(lp):
        return -np.inf
    with np.errstate(all="ignore"):
        ll = normal_loglike(flux, model_flux(x, comps, voigt, lsf, multiplet), sd=sd, noise=noise)
    return lp + ll if np.isfinite(ll) else -np.inf


# --------------------------------------------------------------------------- information criteria
def n_free(n, voigt, sd_free=True):
    """``MAP.len``: total stochastic dimension, ``sd`` included - pinned by
    vamp_1.0/vpfits_intro.ipynb cell 20: AIC - BIC = -95.215 = 17 (2 - ln 2000),
    17 = 4*4 + sd."""
    return (4 if voigt else 3) * n + (1 if sd_free else 0)


def bic(lnL, k, n_data):
    """PyMC 2 ``MAP.BIC = k ln(N) - 2 lnL`` [third party; pinned by the notebook]."""
    return k * np.log(n_data) - 2.0 * lnL


def aic(lnL, k):
    """PyMC 2 ``MAP.AIC = 2 k - 2 lnL``."""
    return 2.0 * k - 2.0 * lnL
