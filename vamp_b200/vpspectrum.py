"""``VPspectrum`` - a whole sightline: detection regions -> region fits -> physical parameters
(the role of vamp_1.0/vpspectrum.py).

Same constructor arguments, attribute names (``region_pixels``, ``region_waves``, ``params``,
``flux_model``, ``regions`` ...) and dictionary keys as the reference class, GPU-shaped inside:

* ``compute_detection_regions`` (vpspectrum.py:67-174): the equivalent widths and the 2 x 9
  Gaussian convolutions run as one batched kernel (K4, ``vamp_detect_host``) - also for many
  sightlines at once (``detection_ratio_batch``); the sequential selection is ``select_regions``.
* ``fit_spectrum`` (vpspectrum.py:243-430): all regions of the sightline (and all candidate
  component counts, restarts and chains of each) are fitted together by ``fit_regions``; the
  reference's retry loop becomes rounds over the regions that have not reached ``chi_limit`` yet.
  N, b, EW and their errors come from ``physics.py`` on the MAP parameters and the device-side
  posterior moments.

The reference file has undefined names on some paths (``attempt_n``, ``n``, ``min_region_size``,
``forced_number_regions`` at vpspectrum.py:203,216,308,317) and asks the MCMC statistics for
``est_sigma_i`` in Voigt mode (KeyError, :401-410); the intended behaviour is implemented here
(Voigt widths are ``GaussianWidth(G_fwhm)``, vpspectrum.py:376-377).  HDF5 I/O goes through ``h5py`` when it is
installed and through ``vamp_b200.h5lite`` (the subset of the format the reference's files use) otherwise;
``set_spectrum`` takes arrays.
"""
import ctypes as C

import numpy as np

from . import _lib, engine
from .fit_regions import estimate_n, estimate_n_doublet, fit_regions
from .physics import (ColumnDensity, DopplerParameter, EquivalentWidthFlux, EquivalentWidthTau, ErrorB, ErrorN,
                      Freq2wave, Tau2flux, Wave2freq)
from .vpfits import VPfit


def detection_ratio_batch(wavelengths, fluxes, noises, std_min=2, std_max=11, ctx=None):
    """Detection ratio of ``S`` sightlines (sequences of 1-D float64 arrays, ragged lengths allowed)
    in one K4 launch.  Returns a list of arrays (``-inf`` at the end pixels, vpspectrum.py:92-122)."""
    ctx = ctx or engine.default_context()
    ctx.use_current_stream()
    S = len(wavelengths)
    if S == 0:
        return []
    npix = np.array([len(w) for w in wavelengths], dtype=np.int32)
    stride = int(npix.max())
    pack = lambda arrs: np.ascontiguousarray(
        np.stack([np.pad(np.asarray(a, dtype=np.float64), (0, stride - len(a))) for a in arrs]))
    wl, fl, ns = pack(wavelengths), pack(fluxes), pack(noises)
    out = np.empty_like(wl)
    for s0 in range(0, S, 65535):
        s1 = min(S, s0 + 65535)
        _lib.check(ctx.lib.vamp_detect_host(ctx.handle, s1 - s0, stride, _lib.as_int_p(npix[s0:s1]), _lib.as_dbl_p(wl[s0:s1]),
                                            _lib.as_dbl_p(fl[s0:s1]), _lib.as_dbl_p(ns[s0:s1]), int(std_min), int(std_max),
                                            _lib.as_dbl_p(out[s0:s1])))
    return [out[s, : npix[s]].copy() for s in range(S)]


def select_regions(det_ratio, flux, noise, min_region_width=2, N_sigma=4.0, extend=False):
    """The sequential part of ``compute_detection_regions`` (vpspectrum.py:124-174): threshold the
    detection ratio, optionally extend to the continuum, merge with the following region when they
    overlap, keep regions that contain a pixel ``N_sigma`` noise deviations deep, pad by 3 pixels.
    Returns ``[[start, end], ...]`` (pixel indices)."""
    P = len(flux)
    start, endpoints = 0, []
    for i in range(P):
        if start == 0 and det_ratio[i] > N_sigma and flux[i] < 1.0:
            start = i
        elif start != 0 and (det_ratio[i] < N_sigma or flux[i] > 1.0):
            if (i - start) > min_region_width:
                endpoints.append([start, i])
            start = 0
    if extend:
        expanded = []
        for s, e in endpoints:
            while s > 0 and flux[s] < 1.0:
                s -= 1
            while e < P - 1 and flux[e] < 1.0:
                e += 1
            expanded.append([s, e])
    else:
        expanded = endpoints
    pixels, buffer = [], 3
    for i, (s, e) in enumerate(expanded):
        if i < len(expanded) - 1 and e > expanded[i + 1][0]:
            e = expanded[i + 1][1]
        deep = np.nonzero(1.0 - np.asarray(flux[s:e]) > np.abs(noise[s:e]) * N_sigma)[0]
        if deep.size:
            if s >= buffer:
                s -= buffer
            if e < P - buffer:
                e += buffer
            pixels.append([int(s), int(e)])
    return pixels


class VPspectrum():

    def __init__(self, line, spectrum_file=None, out_folder=None, voigt=False, chi_limit=1.5, mcmc_cov=False,
                 get_mcmc_err=True, convergence_attempts=10, chi_sq_maximum=10., max_single_region_components=15,
                 ideal_single_region_components=5, min_region_percentage=2., doublet=None, max_region_pixels=256,
                 **fit_kwargs):
        """vpspectrum.py:23-55.  ``fit_kwargs`` (new): GPU controls forwarded to ``fit_regions``
        (iterations, burn, thin, restarts, chains_per_restart, window, nmax, dtype, seed, lsf ...).
        ``doublet`` (new; BASELINE config C3): a name in ``physics.DOUBLETS`` ("CIV", "MgII") or a pair of
        ``(wavelength, f, Gamma)`` lines, ``line`` being the first: the detection regions of the two lines are
        paired (``pair_doublet_regions``) and every absorber is fitted as a doublet with tied velocity, Doppler
        width and column density; the reported parameters are those of the first line.
        ``max_region_pixels`` (new): detection regions longer than this are split before the fit
        (``split_long_regions``).  256 (default) keeps every region on the one-tile sampler kernels; up to 2048
        (single lines) the region is fitted whole, as the reference does, by the long-region kernels."""
        self.line = line
        self.max_region_pixels = int(max_region_pixels)
        self.doublet = doublet
        self.multiplet = None
        if doublet is not None:
            from .physics import DOUBLETS, doublet_ratios
            l1, l2 = DOUBLETS[doublet] if isinstance(doublet, str) else doublet
            self.multiplet = doublet_ratios(l1, l2)
            self.doublet_wave_ratio = l2[0] / l1[0]
        self.spectrum_file = spectrum_file
        self.out_folder = out_folder
        self.voigt = voigt
        self.chi_limit = chi_limit
        self.mcmc_cov = mcmc_cov
        self.get_mcmc_err = get_mcmc_err
        self.convergence_attempts = convergence_attempts
        self.chi_sq_maximum = chi_sq_maximum
        self.max_single_region_components = max_single_region_components
        self.ideal_single_region_components = ideal_single_region_components
        self.min_region_percentage = min_region_percentage
        self.fit_kwargs = fit_kwargs
        self.verbose = False
        self.difficult_fit = False
        if spectrum_file is not None:
            self.read_from_spectrum_file()

    # ------------------------------------------------------------------ input
    def set_spectrum(self, wavelength, flux, noise):
        """Arrays instead of a file (what ``read_from_spectrum_file`` fills, vpspectrum.py:58-64)."""
        self.wavelength_array = np.asarray(wavelength, dtype=np.float64)
        self.flux_array = np.asarray(flux, dtype=np.float64)
        self.noise_array = np.asarray(noise, dtype=np.float64)
        if not (len(self.wavelength_array) == len(self.flux_array) == len(self.noise_array)):
            raise ValueError("wavelength / flux / noise lengths differ")
        self.frequency_array = Wave2freq(self.wavelength_array)
        return self

    def read_from_spectrum_file(self):
        """vpspectrum.py:58-64: HDF5 with ``wavelength``, ``noise``, ``flux`` datasets."""
        try:
            import h5py
        except ImportError:                                # no h5py: the built-in reader of the same file subset
            from .h5lite import read_h5
            data = read_h5(self.spectrum_file)
            self.set_spectrum(data['wavelength'], data['flux'], data['noise'])
            return
        with h5py.File(self.spectrum_file, 'r') as data:
            self.set_spectrum(np.array(data['wavelength'][:]), np.array(data['flux'][:]), np.array(data['noise'][:]))

    # ------------------------------------------------------------------ detection
    def compute_detection_regions(self, min_region_width=2, N_sigma=4.0, extend=False, std_min=2, std_max=11):
        """vpspectrum.py:67-174.  Fills ``region_pixels`` / ``region_waves`` (and ``det_ratio``)."""
        self.num_pixels = len(self.wavelength_array)
        self.det_ratio = detection_ratio_batch([self.wavelength_array], [self.flux_array], [self.noise_array],
                                               std_min, std_max)[0]
        self.region_pixels = select_regions(self.det_ratio, self.flux_array, self.noise_array, min_region_width, N_sigma,
                                            extend)
        self.region_waves = [[self.wavelength_array[s], self.wavelength_array[e]] for s, e in self.region_pixels]
        if self.verbose:
            print('Found {} detection regions.'.format(len(self.region_pixels)))

    def split_difficult_region(self):
        """vpspectrum.py:177-241: a single region with more than ``max_single_region_components``
        estimated lines is split at its highest-flux pixels, at least ``min_region_percentage`` of
        the region apart, into ``n // ideal_single_region_components`` contiguous regions."""
        self.difficult_fit = False
        if len(self.region_pixels) != 1:
            return
        start, end = self.region_pixels[0]
        flux = self.flux_array[start:end]
        n = estimate_n(np.flip(flux, 0))
        if n <= self.max_single_region_components:
            return
        self.difficult_fit = True
        forced = max(1, n // self.ideal_single_region_components)
        k = min(len(flux), 10 * forced)
        ind = np.argpartition(flux, -k)[-k:]
        ind_sorted = np.flip(ind[np.argsort(flux[ind])], axis=0)          # highest flux first
        min_size = len(flux) * (self.min_region_percentage / 100.0)
        points = [start, end]
        for i in ind_sorted:
            if len(points) == forced + 1:
                break
            p = start + int(i)
            if all(abs(p - q) >= min_size for q in points):
                points.append(p)
        points.sort()
        self.region_pixels = [[points[i], points[i + 1]] for i in range(len(points) - 1)]
        self.region_waves = [[self.wavelength_array[s], self.wavelength_array[e]] for s, e in self.region_pixels]

    def split_long_regions(self, max_pixels=256, n_sigma=4.0):
        """New (GPU constraint): the sampler kernels hold a whole region in shared memory - one 256-pixel coordinate
        tile in the fast kernels, up to 2048 pixels in the long-region ones (``max_region_pixels``).  A longer
        detection region is cut where the flux comes closest to the continuum (the same choice
        ``split_difficult_region`` makes, vpspectrum.py:200-232): recursively at the highest-flux pixel of the
        middle half, so both parts keep at least a quarter of the pixels.  Absorbers separated by continuum are
        independent, which is what makes regions the unit of the fit in the first place."""
        out = []
        self.cuts_through_absorption = []       # cut pixels whose flux is NOT continuum (a blended / saturated absorber)
        stack = [list(r) for r in reversed(self.region_pixels)]
        while stack:
            s, e = stack.pop()
            if e - s <= max_pixels:
                out.append([s, e])
                continue
            lo, hi = s + (e - s) // 4, e - (e - s) // 4
            cut = lo + int(np.argmax(self.flux_array[lo:hi]))
            if self.flux_array[cut] < 1.0 - n_sigma * self.noise_array[cut]:
                # no continuum pixel in the middle half: the cut goes through absorption, lines straddling it are
                # fitted in both halves and the per-region n / EW / chi^2 differ from an unsplit fit - say so
                import warnings
                self.cuts_through_absorption.append(int(cut))
                warnings.warn("split_long_regions: region [%d, %d) of %d pixels is cut at pixel %d where the flux (%.3f) is "
                              "more than %.1f sigma below the continuum; the two parts are not independent"
                              % (s, e, e - s, cut, self.flux_array[cut], n_sigma))
            stack.append([cut, e])
            stack.append([s, cut])
        if len(out) != len(self.region_pixels):
            self.region_pixels = out
            self.region_waves = [[self.wavelength_array[s], self.wavelength_array[min(e, len(self.wavelength_array) - 1)]]
                                 for s, e in out]

    def pair_doublet_regions(self, max_pixels=256):
        """New (doublets): a detection region of the first line and the region holding its second line - the one
        that contains the first region's centre moved to ``wavelength x lambda_2/lambda_1`` - are fitted as ONE
        region, from the start of the first to the end of the second (both lines of an absorber must be under the
        same model).  Chains of overlapping pairs merge; a joint region longer than ``max_pixels`` (one coordinate
        tile of the sampler kernels) is left unpaired."""
        if self.multiplet is None or not self.region_pixels:
            return
        regs = sorted([list(r) for r in self.region_pixels])
        wave = self.wavelength_array
        P = len(wave)
        group = list(range(len(regs)))

        def find(i):
            while group[i] != i:
                group[i] = group[group[i]]
                i = group[i]
            return i

        for i, (s, e) in enumerate(regs):
            target = 0.5 * (wave[s] + wave[min(e, P - 1)]) * self.doublet_wave_ratio
            for j, (s2, e2) in enumerate(regs):
                if j != i and wave[s2] <= target <= wave[min(e2, P - 1)]:
                    a, b = find(i), find(j)
                    lo = min(regs[k][0] for k in range(len(regs)) if find(k) in (a, b))
                    hi = max(regs[k][1] for k in range(len(regs)) if find(k) in (a, b))
                    if hi - lo <= max_pixels:
                        group[b] = a
        merged = {}
        for k, (s, e) in enumerate(regs):
            g = find(k)
            lo, hi = merged.get(g, (s, e))
            merged[g] = (min(lo, s), max(hi, e))
        out = sorted([list(v) for v in merged.values()])
        # regions swallowed by a joint region (between the two lines) join it
        final = []
        for s, e in out:
            if final and s < final[-1][1] and max(e, final[-1][1]) - final[-1][0] <= max_pixels:
                final[-1][1] = max(final[-1][1], e)
            else:
                final.append([s, e])
        self.region_pixels = final
        self.region_waves = [[wave[s], wave[min(e, P - 1)]] for s, e in final]

    # ------------------------------------------------------------------ fit
    def _region_arrays(self, start, end):
        """vpspectrum.py:274-279: flipped so that frequency ascends."""
        return (np.flip(self.frequency_array[start:end], 0), np.flip(self.flux_array[start:end], 0),
                np.flip(self.noise_array[start:end], 0), np.flip(self.wavelength_array[start:end], 0))

    def fit_spectrum(self):
        """vpspectrum.py:243-430.  Returns ``params`` (``b, b_std, N, N_std, EW, centers,
        region_numbers``); ``flux_model`` holds the model flux and the per-region fits."""
        if not hasattr(self, "region_pixels"):
            self.compute_detection_regions(min_region_width=2)
            self.split_difficult_region()
        self.pair_doublet_regions()
        self.split_long_regions(self.max_region_pixels if self.multiplet is None else min(self.max_region_pixels, 256))
        R = len(self.region_pixels)
        self.params = {'b': np.array([]), 'b_std': np.array([]), 'N': np.array([]), 'N_std': np.array([]),
                       'EW': np.array([]), 'centers': np.array([]), 'region_numbers': np.array([])}
        self.flux_model = {'total': np.ones(len(self.flux_array)), 'chi_squared': np.zeros(R),
                           'region_pixels': self.region_pixels, 'amplitude': np.array([]), 'sigmas': np.array([]),
                           'centers': np.array([]), 'region_numbers': np.array([]), 'EW': np.zeros(R),
                           'std_a': np.array([]), 'std_s': np.array([]), 'std_c': np.array([]), 'cov_as': np.array([]),
                           'difficult_fit': self.difficult_fit}
        if getattr(self, "cuts_through_absorption", None):               # new key: only present when it applies
            self.flux_model['cuts_through_absorption'] = np.array(self.cuts_through_absorption)
        self.regions = []
        if R == 0:
            return self.params
        arrays = [self._region_arrays(s, e) for s, e in self.region_pixels]
        host = [(a[0], a[1], a[2]) for a in arrays]
        n_est = [estimate_n(a[1]) for a in arrays]
        if self.multiplet is not None:                                    # two significant minima per absorber
            n_est = [estimate_n_doublet(a[1], a[2]) for a in arrays]
        # ---- rounds over the regions that have not converged (the reference's per-region attempt loop)
        attempts_left = []
        for n in n_est:
            if n > self.max_single_region_components:                     # vpspectrum.py:289-295
                self.flux_model['difficult_fit'] = True
                attempts_left.append(1 if n > 1.5 * self.max_single_region_components else 2)
            else:
                attempts_left.append(self.convergence_attempts)
        best = [None] * R
        best_chi = [-1.0] * R
        history = [[] for _ in range(R)]                                  # (n fitted, reduced chi^2) per attempt
        n0 = list(n_est)
        kw = dict(self.fit_kwargs)
        if self.multiplet is not None:
            kw["multiplet"] = self.multiplet
        base_seed = int(kw.pop("seed", 0))
        rnd = 0
        while True:
            todo = [i for i in range(R) if attempts_left[i] > 0]
            if not todo:
                break
            for i in todo:                                                # force more lines after repeated failures
                same = [c for n, c in history[i] if n == n0[i]]          # (vpspectrum.py:306-322)
                if len(same) > 2 and min(same) > self.chi_sq_maximum:
                    n0[i] += 1
            res = fit_regions([host[i] for i in todo], uids=[(rnd << 20) | i for i in todo], voigt=self.voigt,
                              n0=[n0[i] for i in todo], seed=base_seed + rnd, distributed=False,
                              chi_limit=self.chi_limit, **kw)
            for i, r in zip(todo, res):
                P = len(host[i][1])
                freedom = P - 3 * r.n                                     # region.set_freedom() after the fit
                chi = float(r.chi2) / freedom if freedom > 0 else float("inf")
                history[i].append((r.n, chi))
                if best[i] is None or chi < best_chi[i]:
                    best[i], best_chi[i] = r, chi
                attempts_left[i] -= 1
                limit = self.chi_limit * (3.0 if n0[i] > self.max_single_region_components else 1.0)
                if best_chi[i] < limit:
                    attempts_left[i] = 0
            rnd += 1
        # ---- harvest (vpspectrum.py:352-426)
        for j, ((start, end), (nu, fluxes, noise, waves), r) in enumerate(zip(self.region_pixels, arrays, best)):
            n = r.n
            th = r.theta                                                  # [n, 4] (amplitude, centroid, w1, w2) in Hz
            profiles = []
            for k in range(n):
                if self.voigt:
                    profiles.append(VPfit.VoigtFunction(nu, th[k, 1], th[k, 0], th[k, 2], th[k, 3]))
                else:
                    profiles.append(VPfit.GaussFunction(nu, th[k, 0], th[k, 1], th[k, 2]))
                if self.multiplet is not None:                            # the absorber's tied second line
                    fr, ar, lr = self.multiplet
                    if self.voigt:
                        profiles[-1] = profiles[-1] + VPfit.VoigtFunction(nu, fr * th[k, 1], ar * th[k, 0], lr * th[k, 2],
                                                                          fr * th[k, 3])
                    else:
                        profiles[-1] = profiles[-1] + VPfit.GaussFunction(nu, ar * th[k, 0], fr * th[k, 1], fr * th[k, 2])
            total = Tau2flux(np.sum(profiles, axis=0))
            self.flux_model['chi_squared'][j] = best_chi[j]
            self.flux_model['total'][start:end] = np.flip(total, 0)
            self.flux_model['region_' + str(j) + '_wave'] = np.flip(waves, 0)
            self.flux_model['region_' + str(j) + '_flux'] = np.array([np.flip(Tau2flux(p), 0) for p in profiles])
            self.flux_model['EW'][j] = EquivalentWidthFlux(edges=self.flux_model['region_' + str(j) + '_wave'],
                                                           fluxes=self.flux_model['region_' + str(j) + '_flux'])
            heights = th[:, 0].copy()
            centers = Freq2wave(th[:, 1])
            region_numbers = np.arange(n)
            sigmas = VPfit.GaussianWidth(th[:, 3]) if self.voigt else th[:, 2].copy()
            std_a = r.std[:, 0]
            std_s = VPfit.GaussianWidth(r.std[:, 3]) if self.voigt else r.std[:, 2]
            std_c = r.std[:, 1]
            for key, val in (('amplitude', heights), ('centers', centers), ('region_numbers', region_numbers),
                             ('sigmas', sigmas)):
                self.flux_model[key] = np.append(self.flux_model[key], val)
            if self.mcmc_cov:                                             # vpspectrum.py:384-397
                # the reference takes sqrt(diag(np.cov(...))) here (chain_covariance, N - 1 normalised), while the
                # record's ``std`` is PyMC's trace.std() (population): rescale so std_* agree with cov_as
                f = np.sqrt(r.ntally / (r.ntally - 1.0)) if r.ntally > 1 else 1.0
                std_a, std_s, std_c = std_a * f, std_s * f, std_c * f
                for key, val in (('std_a', std_a), ('std_s', std_s), ('std_c', std_c), ('cov_as', r.cov_aw)):
                    self.flux_model[key] = np.append(self.flux_model[key], val)
                self.params['N_std'] = np.append(self.params['N_std'], ErrorN(heights, sigmas, std_a, std_s, r.cov_aw))
            elif self.get_mcmc_err:                                       # vpspectrum.py:399-418
                self.flux_model['std_s'] = np.append(self.flux_model['std_s'], std_s)
                self.flux_model['std_a'] = np.append(self.flux_model['std_a'], std_a)
                self.params['b_std'] = np.append(self.params['b_std'], ErrorB(std_s, self.line))
                self.params['N_std'] = np.append(self.params['N_std'],
                                                 ErrorN(amplitude=heights, sigma=sigmas, std_a=std_a, std_s=std_s,
                                                        cov_as=np.zeros(n)))
            self.params['b'] = np.append(self.params['b'], DopplerParameter(sigmas, self.line))
            self.params['N'] = np.append(self.params['N'], ColumnDensity(heights, sigmas))
            self.params['centers'] = np.append(self.params['centers'], centers)
            self.params['region_numbers'] = np.append(self.params['region_numbers'], region_numbers)
            for k in range(n):
                self.params['EW'] = np.append(self.params['EW'], EquivalentWidthTau(profiles[k], [waves[0], waves[-1]]))
            r.attempt_n = [h[0] for h in history[j]]
            r.attempt_chi_squareds = [h[1] for h in history[j]]
            r.best_chi_squared = best_chi[j]
            self.regions.append(r)
        if self.out_folder is not None:
            name = (self.spectrum_file or "spectrum").split('/', -1)[-1]
            name = name[:name.find('.')] if '.' in name else name
            self.output_filename = self.out_folder + name + ('_voigt_' if self.voigt else '_gauss_')
            self.write_file()
        return self.params

    # ------------------------------------------------------------------ output
    def write_file(self):
        """vpspectrum.py:528-538: ``*params.h5`` and ``*flux_model.h5`` with one dataset per key - through h5py when
        it is installed, otherwise through the built-in writer of the same file subset (``h5lite``)."""
        try:
            import h5py
        except ImportError:
            import os
            from .h5lite import read_h5, write_h5
            for suffix, table in (('params.h5', self.params), ('flux_model.h5', self.flux_model)):
                name = self.output_filename + suffix
                merged = read_h5(name) if os.path.exists(name) else {}      # the reference opens in append mode
                merged.update({k: np.array(v) for k, v in table.items()})
                write_h5(name, merged)
            return
        with h5py.File(self.output_filename + 'params.h5', 'a') as f:
            for p in self.params.keys():
                f.create_dataset(p, data=np.array(self.params[p]))
        with h5py.File(self.output_filename + 'flux_model.h5', 'a') as f:
            for p in self.flux_model.keys():
                f.create_dataset(p, data=np.array(self.flux_model[p]))
