"""``VPfit`` - the reference's fit engine API (vamp_1.0/vpfits.py) on the B200 kernels.

Same class name, method names, argument meaning and read-back attributes as the reference's
``vpfits.VPfit`` so that ``VPregion.region_fit`` / ``VPspectrum`` and the
notebooks keep working; what changed is what runs underneath:

=====================================  ==========================================================
reference (PyMC 2 / astropy, 1 thread)  here (libvamp_b200.so on one B200)
=====================================  ==========================================================
``mc.MCMC(model).sample`` (vpfits:393)  K2: ``n_chains`` warp-resident Metropolis chains, one launch
``mc.MAP(model).fit``     (vpfits:358)  K3: batched Nelder-Mead from every chain's best state
``Voigt1D``/``Tau2flux``/``mc.Normal``  K1 device functions fused into K2/K3
``find_bic`` 3 sequential restarts      3 restart *groups* of ``n_chains/3`` chains, concurrently
=====================================  ==========================================================

Deliberate, documented deviations: many chains instead of one; the MAP polish
starts from each chain's best state instead of its last; after ``find_bic`` the ``.value`` s are
the MAP of the best restart group (the reference keeps restart #3, whatever its quality);
proposal scales are tuned every ``tune_interval`` = 50 iterations (PyMC 2: 1000).
There is no CPU fallback: every method needs the built library and a GPU.
"""
import datetime
import itertools

import numpy as np

from . import engine
from .physics import Tau2flux  # noqa: F401  (re-exported like the reference's `from physics import *`)

_fit_counter = itertools.count(1)


class _Node:
    """Stand-in for a PyMC 2 variable: ``.value`` (lazy), ``.__name__``."""

    def __init__(self, name, getter):
        self.__name__ = name
        self._getter = getter

    @property
    def value(self):
        return self._getter()

    def __array__(self, dtype=None, copy=None):
        # numpy arithmetic on a node uses its current value (PyMC 2 variables overload the operators; the intro
        # notebook passes ``vpfit.std_deviation`` straight into ``VPfit.Chisquared``)
        return np.asarray(self.value, dtype=dtype)


class _MAP:
    """``mc.MAP`` read-backs used by the reference and notebooks: BIC, AIC, lnL, logp."""

    def __init__(self, fit):
        self._fit = fit
        self.BIC = self.AIC = self.lnL = self.logp_at_max = None
        self.len = fit._k
        self.data_len = len(fit._flux)

    def fit(self, iterlim=1000, tol=1e-4, **_):
        self._fit._map_fit(iterlim, tol, store=self)


class _Trace:
    def __init__(self, arr):
        self._a = arr

    def __getitem__(self, idx):
        return self._a[idx]

    def __call__(self):
        return self._a

    def __len__(self):
        return len(self._a)

    def gettrace(self):
        return self._a


class _MCMC:
    """``mc.MCMC`` read-backs: ``trace(name)[:]``, ``stats()``, ``DIC``, ``BPIC``."""

    def __init__(self, fit):
        self._fit = fit

    def sample(self, iter, burn=0, thin=1, **_):
        self._fit._sample(iter, burn, thin)

    def trace(self, name, chain=None):
        name = getattr(name, "__name__", name)
        return _Trace(self._fit._trace(name, chain))

    def stats(self, chain=None):
        return self._fit._stats(chain)

    @property
    def DIC(self):
        return self._fit._dic()[0]

    @property
    def BPIC(self):
        return self._fit._dic()[1]


class VPfit():

    def __init__(self, noise=None, n_chains=96, dtype="f32", seed=None, lsf=None, tune_interval=50,
                 tune_throughout=True, polish_top=4, ctx=None, multiplet=None, step_method="metropolis"):
        """``noise=None``: the noise level ``sd ~ U(0,1)`` is a free parameter (vpfits.py:35-39);
        otherwise the per-pixel noise array.  The keyword arguments are new (GPU controls).
        ``multiplet = (freq_ratio, amp_ratio[, lorentz_ratio])`` (``physics.doublet_ratios``): every component
        is a doublet with tied velocity / width / column density - new, the reference fits single lines."""
        self.multiplet = None if multiplet is None else tuple(float(v) for v in multiplet)
        # "metropolis": the scalar steppers PyMC 2 auto-assigns (what the reference runs); "adaptive": PyMC 2's
        # AdaptiveMetropolis per component, the method mcmc_fit names (vpfits.py:361) - engine.Sampler
        self.step_method = step_method
        self.noise = None if noise is None else np.asarray(noise, dtype=np.float64)
        self.verbose = False
        self.n_chains = max(3, int(n_chains) // 3 * 3)
        self.dtype = dtype
        self.lsf = lsf
        self.tune_interval = int(tune_interval)
        self.tune_throughout = bool(tune_throughout)
        self.polish_top = int(polish_top)
        self._ctx = ctx
        self._seed = int(np.random.SeedSequence().entropy & 0xFFFFFFFFFFFFFFFF) if seed is None else int(seed)
        self._sd = None
        # vpfits.py:39: the Normal's PRECISION, 1 / sd^2 with sd ~ U(0, 1) free, or 1 / noise^2 (the reference keeps the
        # name ``std_deviation`` for it); ``.value`` follows the fit's current sd
        self.std_deviation = _Node("std_deviation", lambda: (1.0 / self.noise ** 2) if self.noise is not None else
                                   (1.0 / self._sd ** 2 if self._sd else np.nan))

    # ------------------------------------------------------------------ static helpers (API surface)
    @staticmethod
    def GaussFunction(x, amplitude, centroid, sigma):
        """vpfits.py:44-54 (also the detection filter of vpspectrum.py:108): host numpy."""
        return amplitude * np.exp(-0.5 * ((x - centroid) / sigma) ** 2)

    @staticmethod
    def VoigtFunction(x, centroid, amplitude, L_fwhm, G_fwhm):
        """vpfits.py:58-76: astropy ``Voigt1D(x_0, amplitude_L, fwhm_L, fwhm_G)(x)``, evaluated in
        float64 by the device Faddeeva (``vamp_profile_host``)."""
        return engine.profile_tau(x, centroid, amplitude, L_fwhm, G_fwhm, voigt=True)

    @staticmethod
    def GaussianWidth(G_fwhm):
        """vpfits.py:80-88."""
        return G_fwhm / (2. * np.sqrt(2. * np.log(2.)))

    @staticmethod
    def Chisquared(observed, expected, noise):
        """vpfits.py:109-118."""
        return float(np.sum(((np.asarray(observed) - np.asarray(expected)) / np.asarray(noise)) ** 2))

    @staticmethod
    def ReducedChisquared(observed, expected, noise, freedom):
        """vpfits.py:121-131."""
        return VPfit.Chisquared(observed, expected, noise) / freedom

    # ------------------------------------------------------------------ model
    def _context(self):
        if self._ctx is None:
            self._ctx = engine.default_context()
        return self._ctx

    def initialise_components(self, frequency_array, n, sigma_max):
        """vpfits.py:219-262.  Gaussian components ``(amplitude, centroid, sigma)``."""
        self._build_nodes(n, voigt=False)

    def initialise_voigt_profiles(self, frequency_array, n, sigma_max, local_minima=[]):
        """vpfits.py:265-307.  Voigt components ``(amplitude, centroid, L_fwhm, G_fwhm)``."""
        self._build_nodes(n, voigt=True)

    def _build_nodes(self, n, voigt):
        self.estimated_variables = {}
        self.estimated_profiles = []
        names = ("amplitude", "centroid", "L_fwhm", "G_fwhm") if voigt else ("amplitude", "centroid", "sigma")
        pymc_names = ("xexp_%d", "est_centroid_%d", "est_L_%d", "est_G_%d") if voigt else \
                     ("xexp_%d", "est_centroid_%d", "est_sigma_%d")
        self._param_index = {}
        for c in range(n):
            self.estimated_variables[c] = {}
            for i, (nm, pn) in enumerate(zip(names, pymc_names)):
                self.estimated_variables[c][nm] = _Node(pn % c, lambda c=c, i=i: float(self._theta[c, i]))
                self._param_index[pn % c] = (c, i)
            self.estimated_profiles.append(_Node("component_%d" % c, lambda c=c: self._component_tau(self._theta[c])))
        self._param_index["sd"] = ("sd", 0)

    def initialise_model(self, frequency, flux, n, local_minima=[], voigt=False, noise_array=None):
        """vpfits.py:310-349: (re)build the model for ``n`` components and draw fresh initial
        values (PyMC 2: ``xexp`` = 0.5, Uniforms at a prior draw) for every chain."""
        frequency = np.asarray(frequency, dtype=np.float64)
        flux = np.asarray(flux, dtype=np.float64)
        if n < 1:
            raise ValueError("n must be >= 1")
        self.sigma_max = (frequency[-1] - frequency[0]) / 2.
        self._voigt = bool(voigt)
        if voigt:
            if self.verbose:
                print("Initialising Voigt profile components.")
            self.fwhm_max = self.sigma_max * 2 * np.sqrt(2 * np.log(2.))
            self.initialise_voigt_profiles(frequency, n, self.fwhm_max, local_minima=local_minima)
        else:
            if self.verbose:
                print("Initialising Gaussian profile components.")
            self.initialise_components(frequency, n, self.sigma_max)
        self._x, self._flux, self._n = frequency, flux, int(n)
        self._sd_free = self.noise is None
        self._k = (4 if voigt else 3) * self._n + (1 if self._sd_free else 0)
        chi_noise = self.noise if self.noise is not None else (
            np.ones_like(flux) if noise_array is None else np.asarray(noise_array, dtype=np.float64))
        self._chi_noise = chi_noise
        ctx = self._context()
        self._regions = engine.Regions(ctx, [(frequency, flux, chi_noise)], lsf=self.lsf, multiplet=self.multiplet)
        B = self.n_chains
        fit_id = next(_fit_counter)
        uid = (np.uint64(fit_id) << np.uint64(24)) + np.arange(B, dtype=np.uint64)
        self._sampler = engine.Sampler(
            self._regions, np.zeros(B, np.int32), np.full(B, self._n, np.int32), uid, voigt=voigt,
            sd_free=self._sd_free, dtype=self.dtype, seed=self._seed, tune_interval=self.tune_interval,
            tune_throughout=self.tune_throughout, step_method=self.step_method)
        self._sampler.init()
        self._primary = 0
        self._ridx = 0
        self._pull_values(self._sampler.get("theta")[self._primary])
        self.total = _Node("profile", self._total_flux)
        self.profile = _Node("obs", lambda: self._flux)
        self.model = self
        self._mcmc_done = False

    # ------------------------------------------------------------------ value plumbing
    def _pull_values(self, theta_row_scaled):
        cm = self._sampler.cmax
        comps = np.asarray(theta_row_scaled[: 4 * cm]).reshape(1, cm, 4)[:, : self._n]
        self._theta = self._regions.to_physical(comps, [self._ridx])[0]
        self._sd = float(theta_row_scaled[4 * cm]) if self._sd_free else None

    def _component_tau(self, row):
        tau = engine.profile_tau(self._x, row[1], row[0], row[2], row[3], voigt=self._voigt, ctx=self._context())
        if self.multiplet is not None and self.multiplet[1] != 0.0:      # the tied second line of the doublet
            fr, ar = self.multiplet[0], self.multiplet[1]
            lr = self.multiplet[2] if len(self.multiplet) > 2 else 1.0
            w1, w2 = (lr * row[2], fr * row[3]) if self._voigt else (fr * row[2], row[3])
            tau = tau + engine.profile_tau(self._x, fr * row[1], ar * row[0], w1, w2, voigt=self._voigt, ctx=self._context())
        return tau

    def _total_flux(self):
        _, _, flux = engine.loglike(self._regions, self._theta[None], [self._ridx], sd=None, voigt=self._voigt,
                                    dtype="f64", want_flux=True)
        return flux[0, : len(self._x)]

    # ------------------------------------------------------------------ MAP
    def map_estimate(self, iterations=2000):
        """vpfits.py:352-358: ``self.map = mc.MAP(self.model); self.map.fit(iterlim, tol=1e-3)``."""
        self.map = _MAP(self)
        self.map.fit(iterlim=iterations, tol=1e-3)

    def _map_fit(self, iterlim, tol, store, groups=None):
        """Nelder-Mead polish (K3) of the best chains; sets the values to the best optimum."""
        s = self._sampler
        cur_logp = s.get("best_logp") if self._mcmc_done else None
        starts = s.get("best_theta") if self._mcmc_done else s.get("theta")
        B = s.B
        groups = groups or [np.arange(B)]
        sel = []
        for g in groups:
            if cur_logp is None:
                sel.append(g[: max(1, min(len(g), self.polish_top * 4))])
            else:
                order = g[np.argsort(-cur_logp[g])]
                sel.append(order[: max(1, min(len(g), self.polish_top))])
        idx = np.concatenate(sel)
        res = s.map_polish(starts[idx], np.zeros(len(idx), np.int32), np.full(len(idx), self._n, np.int32),
                           maxiter=iterlim, xtol=1e-4, ftol=tol, dtype="f64")
        logp = res["logp"].cpu().numpy()
        out, off = [], 0
        for g_sel in sel:
            j = off + int(np.argmax(logp[off: off + len(g_sel)]))
            out.append(j)
            off += len(g_sel)
        self._map_res = {k: v.cpu().numpy() for k, v in res.items()}
        self._map_sel = out
        best = out[int(np.argmax(logp[out]))]
        self._apply_map(best, store)
        return out

    def _apply_map(self, j, store):
        r = self._map_res
        self._pull_values(r["theta"][j])
        store.lnL = float(r["lnL"][j])
        store.logp_at_max = float(r["logp"][j])
        store.BIC = float(r["bic"][j])
        store.AIC = float(2.0 * self._k - 2.0 * r["lnL"][j])
        store.chi2 = float(r["chi2"][j])
        store.iterations = int(r["iters"][j])

    # ------------------------------------------------------------------ MCMC
    def mcmc_fit(self, iterations=15000, burnin=100, thinning=15, step_method=None):
        """vpfits.py:361-395.  ``step_method`` is accepted for signature compatibility (the reference's default,
        ``mc.AdaptiveMetropolis``, is only printed there, never installed: PyMC 2 runs its scalar Metropolis
        steppers); which sampler runs here is the constructor's ``VPfit(step_method="metropolis" | "adaptive")``."""
        try:
            getattr(self, 'map')
        except AttributeError:
            print("\nWARNING: MAP estimate not provided. \nIt is recommended to compute this \
            in advance of running the MCMC so as to start the sampling with good initial values.")
        self.mcmc = _MCMC(self)
        starttime = datetime.datetime.now()
        self.mcmc.sample(iter=iterations, burn=burnin, thin=thinning)
        self._context().synchronize()
        self.fit_time = str(datetime.datetime.now() - starttime)
        print("\nTook:", self.fit_time, " to finish.")

    def _sample(self, iters, burn, thin, keep_trace=True):
        s = self._sampler
        torch = engine._torch()
        if hasattr(self, "map") and self.map.BIC is not None and not self._mcmc_done:
            # like the reference, sampling starts from the MAP values: restart every chain there
            cm = s.cmax
            row = np.zeros(s.ts)
            row[: 4 * self._n] = self._regions.to_scaled(self._theta[None], [0])[0].reshape(-1)
            if self._sd_free:
                row[4 * cm] = self._sd
            s.init(np.tile(row, (s.B, 1)))
        s.burn, s.thin = int(burn), max(1, int(thin))
        cap = max(0, (int(iters) - int(burn) + s.thin - 1) // s.thin + 1) if keep_trace else 0
        s.trace_cap = cap
        s.t["trace"] = torch.zeros((s.B, cap, s.ts), dtype=torch.float64, device=s.ctx.torch_device) if cap else None
        for k in ("mean", "m2", "comom", "sum_lnL", "sum_lnL2"):
            s.t[k].zero_()
        s.t["ntally"].zero_()
        # burn / thin are counted from the start of THIS run, so the iteration counter restarts; the Philox stream must
        # not: every further run of the same sampler gets its own key (the first keeps the fit's seed)
        run = getattr(s, "_runs", 0)
        s.seed = (self._seed + run * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        s._runs = run + 1
        s.iteration = 0
        s.run(int(iters))
        self._mcmc_done = True
        lp = s.get("logp")
        self._primary = int(np.argmax(s.get("best_logp"))) if np.all(np.isfinite(lp)) else 0
        self._pull_values(s.get("theta")[self._primary])

    def _trace(self, name, chain=None):
        s = self._sampler
        if s.t["trace"] is None:
            raise KeyError("no trace was kept")
        ntal = s.get("ntally")
        tr = s.t["trace"].cpu().numpy()
        chains = range(s.B) if chain == "all" else [self._primary if chain is None else int(chain)]
        out = []
        for b in chains:
            T = int(min(ntal[b], s.trace_cap))
            rows = tr[b, :T]
            if name == "sd":
                out.append(rows[:, 4 * s.cmax].copy())
                continue
            if name.startswith("component_"):
                c = int(name.split("_")[1])
                comps = self._regions.to_physical(rows[:, : 4 * s.cmax].reshape(T, s.cmax, 4), np.full(T, self._ridx))
                out.append(np.array([self._component_tau(comps[t, c]) for t in range(T)]))
                continue
            c, i = self._param_index[name]
            comps = self._regions.to_physical(rows[:, : 4 * s.cmax].reshape(T, s.cmax, 4), np.full(T, self._ridx))
            out.append(comps[:, c, i].copy())
        return np.array(out) if chain == "all" else out[0]

    def _stats(self, chain=None):
        """``mcmc.stats()``: per-variable mean / standard deviation / n of the tallied samples
        (device Welford accumulators), keyed by the PyMC 2 variable names."""
        s = self._sampler
        b = self._primary if chain is None else int(chain)
        n = int(s.get("ntally")[b])
        mean, m2 = s.get("mean")[b], s.get("m2")[b]
        scale, nu0 = self._regions.scale[self._ridx], self._regions.nu0[self._ridx]
        out = {}
        for name, (c, i) in self._param_index.items():
            if c == "sd":
                if not self._sd_free:
                    continue
                idx, f, o = 4 * s.cmax, 1.0, 0.0
            else:
                idx = 4 * c + i
                f, o = (1.0, 0.0) if i == 0 else ((scale, nu0) if i == 1 else (scale, 0.0))
            var = m2[idx] / n if n > 0 else np.nan      # PyMC 2 trace.std(): population std
            out[name] = {"n": n, "mean": mean[idx] * f + o, "standard deviation": np.sqrt(var) * f}
        return out

    def gelman_rubin(self, good_only=True):
        """New: multi-chain R-hat per variable (PyMC 2 names) over the chains of this fit, straight from the
        device accumulators (``vamp_b200.diagnostics.rhat``).  ``good_only``: chains within 30 log-posterior units
        of the best one (chains that never left a noise mode would dominate the between-chain variance)."""
        from . import diagnostics
        s = self._sampler
        mean, m2, nt, blp = s.get("mean"), s.get("m2"), s.get("ntally"), s.get("best_logp")
        keep = np.isfinite(blp)
        if good_only:
            keep &= blp > np.nanmax(blp) - 30.0
        r = diagnostics.rhat(mean[keep], m2[keep], nt[keep])
        out = {}
        for name, (c, i) in self._param_index.items():
            if c == "sd":
                if self._sd_free:
                    out[name] = float(r[4 * s.cmax])
            else:
                out[name] = float(r[4 * c + i])
        return out

    def raftery_lewis(self, q=0.995, r=0.05, s=0.95, epsilon=0.001, chain=None):
        """``mc.raftery_lewis(vpfit.mcmc, q, r)`` of the notebooks (vpfits_intro.ipynb cell 18): per variable (PyMC 2
        names) the tuple ``(nmin, kthin, nburn, nprec, kmind)`` from the kept trace of one chain."""
        from . import diagnostics
        out = {}
        for name, (c, _) in self._param_index.items():
            if c == "sd" and not self._sd_free:
                continue
            out[name] = diagnostics.raftery_lewis(self._trace(name, chain), q, r, s, epsilon)
        return out

    def geweke(self, first=0.1, last=0.5, intervals=20, chain=None):
        """``mc.geweke(vpfit.mcmc, intervals=20)`` (vpfits_intro.ipynb cell 17): z-scores per variable."""
        from . import diagnostics
        return {name: diagnostics.geweke(self._trace(name, chain), first, last, intervals)
                for name, (c, _) in self._param_index.items() if not (c == "sd" and not self._sd_free)}

    def _dic(self):
        """PyMC 2: ``DIC = 2 E[D] - D(E[theta])``, ``BPIC = 3 E[D] - 2 D(E[theta])``, D = -2 lnL."""
        s = self._sampler
        b = self._primary
        n = int(s.get("ntally")[b])
        mean_dev = -2.0 * s.get("sum_lnL")[b] / n
        mean = s.get("mean")[b]
        comps = mean[: 4 * s.cmax].reshape(1, s.cmax, 4)[:, : self._n]
        sd = [mean[4 * s.cmax]] if self._sd_free else None
        lnL, _ = engine.loglike(self._regions, comps, [self._ridx], sd=sd, voigt=self._voigt, dtype="f64", scaled=True)
        dev_at_mean = -2.0 * lnL[0]
        return 2 * mean_dev - dev_at_mean, 3 * mean_dev - 2 * dev_at_mean

    # ------------------------------------------------------------------ BIC
    def find_bic(self, frequency_array, flux_array, n, noise_array, freedom, voigt=False,
                 iterations=3000, thin=15, burn=300, thorough=False):
        """vpfits.py:398-429.  Three restarts -> ``bic_array`` and ``red_chi_array`` (length 3).
        Here the restarts are three groups of ``n_chains/3`` chains advanced concurrently
        (``batch.fit_models``): fresh prior-drawn starts, ``iterations`` of MCMC, Nelder-Mead
        polish of each group's best chains, BIC from the likelihood at the group's optimum.
        ``thorough`` doubles the sampling length (the reference runs MAP, MCMC, MAP first)."""
        from . import batch
        frequency_array = np.asarray(frequency_array, dtype=np.float64)
        flux_array = np.asarray(flux_array, dtype=np.float64)
        chi_noise = self.noise if self.noise is not None else np.asarray(noise_array, dtype=np.float64)
        starttime = datetime.datetime.now()
        regions = engine.Regions(self._context(), [(frequency_array, flux_array, chi_noise)], lsf=self.lsf,
                                 multiplet=self.multiplet)
        mb = batch.fit_models(regions, [(0, int(n))], voigt=voigt, sd_free=self.noise is None,
                              iterations=int(iterations) * (2 if thorough else 1), burn=burn, thin=thin, restarts=3,
                              chains_per_restart=self.n_chains // 3, polish_top=self.polish_top, dtype=self.dtype,
                              seed=self._seed, region_uids=[next(_fit_counter)], tune_interval=self.tune_interval,
                              tune_throughout=self.tune_throughout, keep_sampler=True, trace=True,
                              maxiter=iterations, step_method=self.step_method)
        self._adopt(mb, 0, regions, freedom)
        self.fit_time = str(datetime.datetime.now() - starttime)
        return

    def _adopt(self, mb, m, regions, freedom):
        """Take over model ``m`` of a batched fit: values = MAP of the best restart group,
        ``bic_array`` / ``red_chi_array`` in restart order, sampler read-backs of the winning chain."""
        r, n = mb.models[m]
        self._regions, self._ridx = regions, r
        self._voigt, self._n, self._sd_free = mb.voigt, n, mb.sd_free
        self._build_nodes(n, mb.voigt)
        self._k = (4 if mb.voigt else 3) * n + (1 if mb.sd_free else 0)
        self._sampler = getattr(mb, "sampler", None)
        self._mcmc_done = True
        P = int(regions.npix[r])
        self._x = None
        self._adopt_arrays(mb, m)
        self.bic_array = [float(v) for v in mb.bic[m]]
        self.red_chi_array = [float(v) / freedom for v in mb.chi2[m]]
        g = int(np.nanargmin(mb.bic[m])) if np.any(np.isfinite(mb.bic[m])) else 0
        self._theta = np.array(mb.theta[m, g, :n])
        self._sd = float(mb.sd[m, g]) if mb.sd_free else None
        self._primary = int(mb.primary_rows[m, g])
        self._post_mean = np.array(mb.mean[m, g, :n])
        self._post_std = np.array(mb.std[m, g, :n])
        self.map = _MAP.__new__(_MAP)
        self.map._fit = self
        self.map.len, self.map.data_len = self._k, P
        self.map.lnL, self.map.logp_at_max = float(mb.lnL[m, g]), float(mb.logp[m, g])
        self.map.BIC = float(mb.bic[m, g])
        self.map.AIC = float(2.0 * self._k - 2.0 * mb.lnL[m, g])
        self.map.chi2 = float(mb.chi2[m, g])
        self.mcmc = _MCMC(self)
        self.sigma_max = (self._x[-1] - self._x[0]) / 2.
        if mb.voigt:
            self.fwhm_max = self.sigma_max * 2 * np.sqrt(2 * np.log(2.))
        self.total = _Node("profile", self._total_flux)
        self.profile = _Node("obs", lambda: self._flux)
        self.model = self

    def _adopt_arrays(self, mb, m):
        src = getattr(mb, "host_regions", None)
        r = mb.models[m][0]
        if src is not None:
            self._x, self._flux, self._chi_noise = (np.asarray(a, dtype=np.float64) for a in src[r])

    def chain_covariance(self, n, voigt=False):
        """vpfits.py:432-456: per component the 3x3 covariance of the (amplitude, sigma, centroid)
        samples (sigma = GaussianWidth(G_fwhm) for Voigt), ``np.cov`` normalisation (N-1).
        From the device co-moment accumulators of the primary chain."""
        s = self._sampler
        b = self._primary
        N = int(s.get("ntally")[b])
        m2, com = s.get("m2")[b], s.get("comom")[b]
        sc = self._regions.scale[self._ridx]
        fw = sc / (2. * np.sqrt(2. * np.log(2.))) if voigt else sc
        wi = 3 if voigt else 2                      # index of G_fwhm / sigma
        pair = {(0, 1): 0, (0, 2): 1, (0, 3): 2, (1, 2): 3, (1, 3): 4, (2, 3): 5}
        cov = np.zeros((n, 3, 3))
        d = float(max(N - 1, 1))
        for i in range(n):
            va, vs, vc = m2[4 * i] / d, m2[4 * i + wi] * fw * fw / d, m2[4 * i + 1] * sc * sc / d
            c_as = com[i, pair[(0, wi)]] * fw / d
            c_ac = com[i, pair[(0, 1)]] * sc / d
            c_sc = com[i, pair[(1, wi)]] * fw * sc / d
            cov[i] = [[va, c_as, c_ac], [c_as, vs, c_sc], [c_ac, c_sc, vc]]
        return cov
