"""``VPregion`` - per-region component-count selection (the role of vamp_1.0/vpregion.py).

Same constructor, attributes (``n``, ``freedom``, ``fit``, ``voigt``, ``chi_limit``...) and
``region_fit`` outcome as the reference class; the implementation is GPU-shaped: instead of
fitting candidate n, n+1, ... one after the other (vpregion.py:58-60), a *window* of
candidates is fitted concurrently by ``batch.fit_models`` and the reference's acceptance rule
(vpregion.py:61-89) is replayed on the resulting BIC / chi^2 tables (``batch.ladder_replay``);
the window is extended only if the rule runs past it.  ``fit_region_pygadds``
(vpregion.py:94-144) is dead py2-only code in the reference and is not provided.
"""
import numpy as np

from . import batch, engine
from .vpfits import VPfit, _fit_counter


class VPregion():

    def __init__(self, frequency_array, flux_array, noise_array, voigt=False, chi_limit=1.5, window=3,
                 **fit_kwargs):
        """vpregion.py:10-19.  New keywords: ``window`` = candidates fitted per launch;
        ``fit_kwargs`` = GPU controls forwarded to ``VPfit`` (n_chains, dtype, seed, lsf...)."""
        self.frequency_array = np.asarray(frequency_array, dtype=np.float64)
        self.flux_array = np.asarray(flux_array, dtype=np.float64)
        self.noise_array = np.asarray(noise_array, dtype=np.float64)
        self.voigt = voigt
        self.chi_limit = chi_limit
        self.num_pixels = len(flux_array)
        self.window = max(1, int(window))
        self.fit_kwargs = fit_kwargs
        self.estimate_n()
        self.set_freedom()

    def estimate_n(self):
        """Initial guess (vpregion.py:21-35): count of local minima of the flux smoothed with a
        sigma = 3 px Gaussian, reset to 1 when below 4."""
        from scipy.ndimage import gaussian_filter
        from scipy.signal import argrelextrema
        minima = argrelextrema(gaussian_filter(self.flux_array, 3), np.less)[0]
        self.n = int(minima.shape[0]) if minima.shape[0] >= 4 else 1

    def set_freedom(self):
        """vpregion.py:37-39: pixels - 3 n (also for Voigt fits; reference quirk kept)."""
        self.freedom = self.num_pixels - 3 * self.n

    def region_fit(self, verbose=True, iterations=3000, thin=15, burn=300):
        """Select the component count by BIC / reduced chi^2 and keep that fit in ``self.fit``.

        Outcome-compatible with vpregion.py:42-91, including its quirks: d.o.f. frozen at the
        initial n, ``thin`` not forwarded (find_bic's default 15 is used), first candidate
        compared through its last restart's BIC."""
        proto = VPfit(**self.fit_kwargs)
        regions = engine.Regions(proto._context(), [(self.frequency_array, self.flux_array, self.noise_array)],
                                 lsf=proto.lsf, multiplet=proto.multiplet)
        uid = next(_fit_counter)
        n0 = self.n
        if verbose:
            print("Setting initial number of lines to: {}".format(n0))
        tables, fits = {}, {}
        max_n = max(1, (self.num_pixels - 1) // 2)
        nxt = n0
        while True:
            cand = [n for n in range(nxt, nxt + self.window) if n <= max_n] or [nxt]
            mb = batch.fit_models(regions, [(0, n) for n in cand], voigt=self.voigt, sd_free=proto.noise is None,
                                  iterations=iterations, burn=burn, thin=15, restarts=3,
                                  chains_per_restart=proto.n_chains // 3, polish_top=proto.polish_top,
                                  dtype=proto.dtype, seed=proto._seed, region_uids=[uid],
                                  tune_interval=proto.tune_interval, tune_throughout=proto.tune_throughout,
                                  keep_sampler=True, trace=True)
            for m, n in enumerate(cand):
                tables[n] = (list(mb.bic[m]), list(mb.chi2[m] / self.freedom))
                fits[n] = (mb, m)
            kept, exit_n, decided = batch.ladder_replay(n0, tables, self.chi_limit, verbose=verbose)
            if decided or cand[-1] >= max_n:
                break
            nxt = cand[-1] + 1
        self.n = exit_n if decided else kept
        self.history = [(n, tables[n][0], tables[n][1]) for n in sorted(tables)]
        mb, m = fits[kept]
        fit = VPfit(**self.fit_kwargs)
        fit._adopt(mb, m, regions, self.freedom)
        self.fit = fit
