"""Oracle: restatement of the PyMC 2.3 machinery the reference delegates to.

TEST INFRASTRUCTURE ONLY.  PyMC 2 (``import pymc as mc``, vamp_1.0/vpfits.py:16;
env name ``vamp27pymc237`` at vpfits.py:345) is a third-party dependency absent
from /root/reference and not installable here (Fortran, Python 2 era), so its
published algorithm is restated from the PyMC 2.3.x sources [3P-memory in
SURVEY.md section 3.2] and anchored on the reference's call sites:

* ``mcmc.sample(iter, burn, thin)``      - vpfits.py:393, 423, 425
* ``mc.MAP(model).fit(iterlim, tol)``    - vpfits.py:357-358, 419-426
* ``map.BIC``                            - vpfits.py:427
* ``find_bic`` 3-restart loop            - vpfits.py:398-429

PARITY UNPINNED: statistical agreement only (no PyMC chains are recorded in the
reference; the only anchors are notebook BIC ladders).
"""
import numpy as np
from scipy.optimize import fmin

from . import density as D


class Model:
    """Parameter-vector view of ``VPfit.initialise_model`` (vpfits.py:310-349).

    Flat vector layout: per component ``[amplitude, centroid, w1(, w2)]`` then
    ``sd`` when ``noise is None`` (``VPfit(noise=None)``, vpfits.py:35-39)."""

    def __init__(self, x, flux, n, voigt, noise=None, lsf=None, multiplet=None):
        self.x = np.asarray(x, dtype=np.float64)
        self.flux = np.asarray(flux, dtype=np.float64)
        self.n, self.voigt, self.noise, self.lsf = int(n), bool(voigt), noise, lsf
        self.multiplet = multiplet            # (freq_ratio, amp_ratio, lorentz_ratio): density.second_line
        self.per = 4 if voigt else 3
        self.sd_free = noise is None
        self.k = self.per * self.n + (1 if self.sd_free else 0)
        self.c_lo, self.c_hi, self.w_max = D.prior_bounds(self.x, voigt)

    def unpack(self, theta):
        comps = np.zeros((self.n, 4))
        comps[:, : self.per] = np.asarray(theta[: self.per * self.n]).reshape(self.n, self.per)
        sd = theta[-1] if self.sd_free else None
        return comps, sd

    def initial(self, rng):
        """PyMC 2 initial values: ``xexp`` value=0.5 (vpfits.py:240, 284); a
        ``mc.Uniform`` without ``value=`` draws from its prior [3P]."""
        theta = np.empty(self.k)
        for c in range(self.n):
            theta[c * self.per] = 0.5
            theta[c * self.per + 1] = rng.uniform(self.c_lo, self.c_hi)
            theta[c * self.per + 2: (c + 1) * self.per] = rng.uniform(0.0, self.w_max, self.per - 2)
        if self.sd_free:
            theta[-1] = rng.uniform(0.0, 1.0)
        return theta

    def logp(self, theta):
        comps, sd = self.unpack(theta)
        return D.log_posterior(comps, self.x, self.flux, self.voigt, sd=sd,
                               noise=None if self.sd_free else self.noise, lsf=self.lsf, multiplet=self.multiplet)

    def loglike(self, theta):
        comps, sd = self.unpack(theta)
        return D.normal_loglike(self.flux, D.model_flux(self.x, comps, self.voigt, self.lsf, self.multiplet), sd=sd,
                                noise=None if self.sd_free else self.noise)

    def total(self, theta):
        comps, _ = self.unpack(theta)
        return D.model_flux(self.x, comps, self.voigt, self.lsf, self.multiplet)


def pymc2_tune(asf, acc_rate):
    """``Metropolis.tune`` switch of PyMC 2.3 [3P]."""
    if acc_rate < 0.001:
        return asf * 0.1
    if acc_rate < 0.05:
        return asf * 0.5
    if acc_rate < 0.2:
        return asf * 0.9
    if acc_rate > 0.95:
        return asf * 10.0
    if acc_rate > 0.75:
        return asf * 2.0
    if acc_rate > 0.5:
        return asf * 1.1
    return asf


def metropolis_sample(model, theta0, iters, burn, thin, rng, tune_interval=1000,
                      tune_throughout=True, proposal_sd=None):
    """``MCMC.sample`` with PyMC 2's auto-assigned scalar ``Metropolis`` steppers
    (the reference names AdaptiveMetropolis but never installs it:
    vpfits.py:382-389 only prints).  One iteration = one Normal random-walk
    step per stochastic, ``proposal_sd = |value0|`` (1 if 0) scaled by the
    adaptive scale factor; tallies when ``i >= burn and i % thin == 0``."""
    theta = np.array(theta0, dtype=np.float64)
    k = model.k
    psd = np.where(theta != 0.0, np.abs(theta), 1.0) if proposal_sd is None else np.asarray(proposal_sd, float)
    asf = np.ones(k)
    acc = np.zeros(k)
    rej = np.zeros(k)
    logp = model.logp(theta)
    trace = []
    for i in range(iters):
        if i and i % tune_interval == 0 and (tune_throughout or i <= burn):
            for j in range(k):
                if acc[j] + rej[j] > 0:
                    asf[j] = pymc2_tune(asf[j], acc[j] / (acc[j] + rej[j]))
            acc[:] = 0
            rej[:] = 0
        for j in range(k):
            old = theta[j]
            theta[j] = rng.normal(old, asf[j] * psd[j])
            logp_p = model.logp(theta)
            if not np.isfinite(logp_p) or np.log(rng.random()) > logp_p - logp:
                theta[j] = old
                rej[j] += 1
            else:
                logp = logp_p
                acc[j] += 1
        if i >= burn and i % thin == 0:
            trace.append(theta.copy())
    return np.array(trace), theta, logp


class CachedModel(Model):
    """``Model`` whose ``logp`` re-evaluates only the component that changed since the last call - the economy PyMC 2's
    LazyFunction cache gives the reference (a step on a parameter of component c recomputes only
    profile c, then ``total`` and the Normal logp).  Same floats as ``Model.logp`` (the per-component rows are summed in
    the same order); it only makes the restated sampler fast enough to run ladders of many-component regions."""

    def __init__(self, *a, **kw):
        Model.__init__(self, *a, **kw)
        self._comps = None
        self._taus = None

    def logp(self, theta):
        comps, sd = self.unpack(theta)
        lp = D.log_prior(comps, self.x, self.voigt, sd=sd)
        if not np.isfinite(lp):
            return -np.inf
        if self._comps is None:
            self._comps = np.full_like(comps, np.nan)
            self._taus = np.zeros((self.n, len(self.x)))
        for c in range(self.n):
            if not np.array_equal(comps[c], self._comps[c]):
                self._taus[c] = D.component_taus(self.x, comps[c: c + 1], self.voigt, self.multiplet)[0]
                self._comps[c] = comps[c]
        with np.errstate(all="ignore"):
            m = D.lsf_convolve(D.tau2flux(self._taus.sum(axis=0)), self.lsf)
            ll = D.normal_loglike(self.flux, m, sd=sd, noise=None if self.sd_free else self.noise)
        return lp + ll if np.isfinite(ll) else -np.inf


def map_fit(model, theta0, iterlim=3000, tol=1e-3):
    """``mc.MAP(model).fit(iterlim=..., tol=1e-3)`` (vpfits.py:358, 422-426):
    ``scipy.optimize.fmin`` (Nelder-Mead) on ``-logp`` with ``ftol=tol``,
    ``maxiter=iterlim`` [3P]; then BIC/AIC from the *likelihood* at the optimum."""
    def neg(p):
        v = model.logp(p)
        return -v if np.isfinite(v) else np.inf
    theta = fmin(neg, np.asarray(theta0, float), ftol=tol, maxiter=iterlim, disp=0)
    lnL = model.loglike(theta)
    return theta, lnL, D.bic(lnL, model.k, len(model.flux)), D.aic(lnL, model.k)


def find_bic(x, flux, n, noise_array, freedom, voigt=False, iterations=3000, thin=15, burn=300,
             rng=None, restarts=3, tune_interval=1000):
    """``VPfit.find_bic`` - vpfits.py:398-429 (non-``thorough`` branch):
    3 x (fresh model, MCMC, MAP from the last MCMC state) -> ``bic_array``,
    ``red_chi_array``; returns also the last restart's MAP vector."""
    rng = rng or np.random.default_rng(0)
    model = CachedModel(x, flux, n, voigt)
    bics, reds, theta = [], [], None
    for _ in range(restarts):
        _, last, _ = metropolis_sample(model, model.initial(rng), iterations, burn, thin, rng,
                                       tune_interval=tune_interval)
        theta, _, b, _ = map_fit(model, last, iterlim=iterations)
        bics.append(b)
        reds.append(D.reduced_chisquared(flux, model.total(theta), noise_array, freedom))
    return bics, reds, theta
