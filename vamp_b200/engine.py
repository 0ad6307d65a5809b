"""Host-side engine: torch tensors for device memory + the C ABI for every kernel.

PyTorch is used only as plumbing (device allocations, streams, ``torch.distributed``); all
arithmetic of the hot path runs in ``libvamp_b200.so``.  There is no CPU fallback: every call
here needs a B200 and the built library.

Units: the public Python API speaks the reference's physical units (Hz).  The kernels work
in *scaled* units (``u = (nu - nu_0)/s`` per region, see include/vamp_b200.h); the
conversions live in :class:`Regions`.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import VAMP_F32, VAMP_F64, VAMP_MIXED, VAMP_GAUSS, VAMP_VOIGT, McmcArgs, MapArgs, check

FWHM_PER_SIGMA = 2.0 * np.sqrt(2.0 * np.log(2.0))
TILE_PIXELS = 256             # one coordinate tile: the regions the one-tile sampler / warp-per-start MAP kernels take
MAX_SAMPLER_PIXELS = 2048     # VAMP_MAX_FIT_PIXELS (include/vamp_b200.h): longest region of the long-region sampler kernels
LONG_CLASS_EDGES = (256, 512, 1024)     # pixel classes of the sampler launches: <= 256 (one tile), <= 512, <= 1024, <= 2048


def long_class(npix):
    """Pixel class of regions of ``npix`` pixels (0: one coordinate tile; 1..3: the long-region kernels' buffer
    lengths).  Chains of different classes go out as separate launches (``batch.fit_models``)."""
    return np.searchsorted(np.asarray(LONG_CLASS_EDGES), np.asarray(npix), side="left")
_DTYPES = {"f32": VAMP_F32, "f64": VAMP_F64, "float32": VAMP_F32, "float64": VAMP_F64, VAMP_F32: VAMP_F32,
           VAMP_F64: VAMP_F64}


def _torch():
    import torch
    return torch


def dtype_code(dtype):
    try:
        return _DTYPES[dtype]
    except KeyError:
        raise ValueError("dtype must be 'f32' or 'f64', got %r" % (dtype,))


def kind_code(voigt):
    return VAMP_VOIGT if voigt else VAMP_GAUSS


class Context:
    """One per process/GPU: wraps ``vamp_ctx`` and follows torch's current stream."""

    def __init__(self, device=None):
        torch = _torch()
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.VampError("no CUDA device visible: vamp_b200 has no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        h = C.c_void_p()
        check(self.lib.vamp_ctx_create(self.device, C.byref(h)))
        self.handle = h
        self.torch_device = torch.device("cuda", self.device)
        self.use_current_stream()

    def use_current_stream(self):
        torch = _torch()
        s = torch.cuda.current_stream(self.device).cuda_stream
        check(self.lib.vamp_ctx_set_stream(self.handle, C.c_void_p(s)))

    def synchronize(self):
        check(self.lib.vamp_ctx_synchronize(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.vamp_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=None):
    torch = _torch()
    if not torch.cuda.is_available():
        raise _lib.VampError("no CUDA device visible: vamp_b200 has no CPU fallback")
    dev = torch.cuda.current_device() if device is None else int(device)
    if dev not in _default_ctx:
        _default_ctx[dev] = Context(dev)
    return _default_ctx[dev]


class Regions:
    """A set of absorption regions resident on the device (``vamp_regions``).

    ``regions``: sequence of ``(nu, flux, noise)`` float64 arrays, ``nu`` ascending (what
    ``VPregion.__init__`` receives, vamp_1.0/vpregion.py:10-19).

    ``multiplet``: ``None`` (single lines, the reference), one ``(freq_ratio, amp_ratio[, lorentz_ratio])``
    tuple for every region, or a per-region list of such tuples / ``None`` - every component of the region then
    carries a second line with tied parameters (``vamp_regions_set_multiplet``, include/vamp_b200.h;
    :func:`vamp_b200.physics.doublet_ratios` gives the ratios of a doublet)."""

    def __init__(self, ctx, regions, lsf=None, multiplet=None):
        self.ctx = ctx
        self.lib = ctx.lib
        regions = list(regions)
        self.host_regions = regions          # kept for read-backs (VPfit._x / _flux)
        self.R = len(regions)
        self.npix = np.array([len(r[0]) for r in regions], dtype=np.int32)
        nu = np.ascontiguousarray(np.concatenate([np.asarray(r[0], dtype=np.float64) for r in regions]))
        flux = np.ascontiguousarray(np.concatenate([np.asarray(r[1], dtype=np.float64) for r in regions]))
        noise = np.ascontiguousarray(np.concatenate([np.asarray(r[2], dtype=np.float64) for r in regions]))
        if not (len(nu) == len(flux) == len(noise) == int(self.npix.sum())):
            raise ValueError("nu / flux / noise lengths differ")
        bad = ~(np.isfinite(flux) & np.isfinite(noise) & (noise != 0.0))
        if bad.any():            # mirrored in vamp_regions_create (error -4): fail before anything is uploaded
            i = int(np.argmax(bad))
            r = int(np.searchsorted(np.cumsum(self.npix), i, side="right"))
            raise ValueError("region %d: flux must be finite and noise finite and non-zero (flux %r, noise %r at pixel %d)"
                             % (r, flux[i], noise[i], i - int(self.npix[:r].sum())))
        self.lsf = None if lsf is None else np.ascontiguousarray(np.asarray(lsf, dtype=np.float64))
        h = C.c_void_p()
        check(self.lib.vamp_regions_create(
            ctx.handle, self.R, _lib.as_int_p(self.npix), _lib.as_dbl_p(nu), _lib.as_dbl_p(flux), _lib.as_dbl_p(noise),
            None if self.lsf is None else _lib.as_dbl_p(self.lsf), 0 if self.lsf is None else int(self.lsf.size),
            C.byref(h)))
        self.handle = h
        self.nu0 = np.empty(self.R)
        self.scale = np.empty(self.R)
        check(self.lib.vamp_regions_scale(h, _lib.as_dbl_p(self.nu0), _lib.as_dbl_p(self.scale)))
        self.max_npix = int(self.npix.max())
        self.h2d_bytes = int(nu.nbytes + flux.nbytes + noise.nbytes + self.npix.nbytes)
        self.multiplet = None
        if multiplet is not None:
            self.set_multiplet(multiplet)

    def set_multiplet(self, multiplet):
        """Install (or, with ``None``, clear) the regions' multiplet table."""
        if multiplet is None:
            check(self.lib.vamp_regions_set_multiplet(self.handle, None, None, None))
            self.multiplet = None
            return
        if len(multiplet) in (2, 3) and np.isscalar(multiplet[0]):
            multiplet = [tuple(multiplet)] * self.R
        if len(multiplet) != self.R:
            raise ValueError("multiplet: one (freq_ratio, amp_ratio[, lorentz_ratio]) per region expected")
        tab = np.zeros((self.R, 3))
        tab[:, 0] = 1.0
        tab[:, 2] = 1.0
        for r, m in enumerate(multiplet):
            if m is not None:
                tab[r, : len(m)] = m
        cols = [np.ascontiguousarray(tab[:, i]) for i in range(3)]
        check(self.lib.vamp_regions_set_multiplet(self.handle, _lib.as_dbl_p(cols[0]), _lib.as_dbl_p(cols[1]),
                                                  _lib.as_dbl_p(cols[2])))
        self.multiplet = tab if np.any(tab[:, 1] != 0.0) else None

    # -- unit conversion ------------------------------------------------------------------
    def to_scaled(self, theta, item_region):
        """theta ``[B, cmax, 4]`` (amplitude, centroid, w1, w2) Hz -> scaled units."""
        th = np.array(theta, dtype=np.float64, copy=True)
        r = np.asarray(item_region)
        s = self.scale[r][:, None]
        th[:, :, 1] = (th[:, :, 1] - self.nu0[r][:, None]) / s
        th[:, :, 2] /= s
        th[:, :, 3] /= s
        return th

    def to_physical(self, theta, item_region):
        th = np.array(theta, dtype=np.float64, copy=True)
        r = np.asarray(item_region)
        s = self.scale[r][:, None]
        th[:, :, 1] = th[:, :, 1] * s + self.nu0[r][:, None]
        th[:, :, 2] *= s
        th[:, :, 3] *= s
        return th

    def close(self):
        if getattr(self, "handle", None):
            self.lib.vamp_regions_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_theta(comps, sd, cmax):
    """``comps [B, n<=cmax, 4]`` + ``sd [B]`` (or None) -> flat ``[B, 4*cmax+1]`` layout."""
    comps = np.asarray(comps, dtype=np.float64)
    B, n = comps.shape[0], comps.shape[1]
    out = np.zeros((B, 4 * cmax + 1))
    out[:, : 4 * n] = comps.reshape(B, 4 * n)
    if sd is not None:
        out[:, 4 * cmax] = sd
    return out


def loglike(regions, comps, item_region, item_ncomp=None, sd=None, voigt=True, dtype="f32", want_flux=False,
            scaled=False):
    """K1 through the device-pointer ABI.  ``comps [B, cmax, 4]`` in Hz (or scaled units when
    ``scaled``).  Returns ``(lnL, chi2[, flux])`` as numpy arrays (flux ``[B, max_npix]``)."""
    torch = _torch()
    ctx = regions.ctx
    ctx.use_current_stream()
    comps = np.asarray(comps, dtype=np.float64)
    B, cmax = comps.shape[0], comps.shape[1]
    item_region = np.asarray(item_region, dtype=np.int32)
    item_ncomp = np.full(B, cmax, dtype=np.int32) if item_ncomp is None else np.asarray(item_ncomp, dtype=np.int32)
    th = comps if scaled else regions.to_scaled(comps, item_region)
    dev = ctx.torch_device
    d_th = torch.from_numpy(pack_theta(th, None, cmax)).to(dev)
    d_ir = torch.from_numpy(item_region).to(dev)
    d_in = torch.from_numpy(item_ncomp).to(dev)
    d_sd = None if sd is None else torch.from_numpy(np.asarray(sd, dtype=np.float64)).to(dev)
    d_lnL = torch.empty(B, dtype=torch.float64, device=dev)
    d_chi = torch.empty(B, dtype=torch.float64, device=dev)
    code = dtype_code(dtype)
    d_flux = None
    if want_flux:
        d_flux = torch.zeros((B, regions.max_npix), dtype=torch.float32 if code == VAMP_F32 else torch.float64, device=dev)
    check(ctx.lib.vamp_loglike_dev(
        ctx.handle, regions.handle, code, kind_code(voigt), B, d_ir.data_ptr(), d_in.data_ptr(), cmax,
        d_th.data_ptr(), 4 * cmax + 1, None if d_sd is None else d_sd.data_ptr(), d_lnL.data_ptr(), d_chi.data_ptr(),
        None if d_flux is None else d_flux.data_ptr(), regions.max_npix))
    out = (d_lnL.cpu().numpy(), d_chi.cpu().numpy())
    if want_flux:
        out = out + (d_flux.double().cpu().numpy(),)
    return out


def loglike_host(regions, comps, item_region, item_ncomp=None, sd=None, voigt=True, dtype="f32", want_flux=False):
    """K1 through the HOST-buffer ABI (``vamp_loglike_host``): what a non-torch binder calls."""
    ctx = regions.ctx
    ctx.use_current_stream()
    comps = np.asarray(comps, dtype=np.float64)
    B, cmax = comps.shape[0], comps.shape[1]
    item_region = np.ascontiguousarray(item_region, dtype=np.int32)
    item_ncomp = np.full(B, cmax, dtype=np.int32) if item_ncomp is None else np.ascontiguousarray(item_ncomp, dtype=np.int32)
    th = np.ascontiguousarray(pack_theta(comps, None, cmax))
    sd_a = None if sd is None else np.ascontiguousarray(sd, dtype=np.float64)
    lnL = np.empty(B)
    chi = np.empty(B)
    flux = np.zeros((B, regions.max_npix)) if want_flux else None
    check(ctx.lib.vamp_loglike_host(
        ctx.handle, regions.handle, dtype_code(dtype), kind_code(voigt), B, _lib.as_int_p(item_region),
        _lib.as_int_p(item_ncomp), cmax, _lib.as_dbl_p(th), 4 * cmax + 1, None if sd_a is None else _lib.as_dbl_p(sd_a),
        _lib.as_dbl_p(lnL), _lib.as_dbl_p(chi), None if flux is None else _lib.as_dbl_p(flux), regions.max_npix))
    return (lnL, chi, flux) if want_flux else (lnL, chi)


def profile_tau(x, centroid, amplitude, w1, w2=0.0, voigt=True, ctx=None):
    """tau(x) of one component, float64 on the device (``vamp_profile_host``)."""
    ctx = ctx or default_context()
    ctx.use_current_stream()
    x = np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)
    out = np.empty_like(x)
    check(ctx.lib.vamp_profile_host(ctx.handle, kind_code(voigt), x.size, _lib.as_dbl_p(x), float(centroid),
                                    float(amplitude), float(w1), float(w2), _lib.as_dbl_p(out)))
    return out


def faddeeva_H(x, y, dtype="f32", ctx=None):
    ctx = ctx or default_context()
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty_like(x)
    check(ctx.lib.vamp_faddeeva_host(ctx.handle, dtype_code(dtype), x.size, _lib.as_dbl_p(x), _lib.as_dbl_p(y),
                                     _lib.as_dbl_p(out)))
    return out


def philox(counter, key, ctx=None):
    ctx = ctx or default_context()
    c = (C.c_uint * 4)(*[int(v) for v in counter])
    k = (C.c_uint * 2)(*[int(v) for v in key])
    o = (C.c_uint * 4)()
    check(ctx.lib.vamp_philox_host(ctx.handle, c, k, o))
    return [int(v) for v in o]


class Sampler:
    """Batched chains on the device (K2) + MAP polish (K3).

    ``item_region[B]``, ``item_ncomp[B]``: region and component count of every chain;
    ``chain_uid[B]``: global 64-bit chain ids (Philox counters - results do not depend on how
    chains are batched or sharded)."""

    def __init__(self, regions, item_region, item_ncomp, chain_uid, voigt=True, sd_free=True, dtype="f32",
                 seed=0, burn=300, thin=15, tune_interval=100, tune_throughout=False, trace_cap=0, cmax=None,
                 step_method="metropolis", am_delay=1000, am_interval=200):
        torch = _torch()
        self.regions = regions
        self.ctx = regions.ctx
        self.lib = self.ctx.lib
        self.voigt = bool(voigt)
        self.kind = kind_code(voigt)
        self.dtype = dtype_code(dtype)
        self.sd_free = bool(sd_free)
        self.B = len(item_region)
        self.item_region_h = np.asarray(item_region, dtype=np.int32)
        self.item_ncomp_h = np.asarray(item_ncomp, dtype=np.int32)
        self.max_ncomp = int(self.item_ncomp_h.max())
        longest = int(regions.npix[self.item_region_h].max()) if self.B else 0
        if longest > MAX_SAMPLER_PIXELS or (longest > TILE_PIXELS and (regions.multiplet is not None or step_method == "adaptive")):
            # the C ABI answers -3 at launch; say it here, before any state is allocated, with the way out
            raise _lib.VampError(
                "the sampler holds a whole region in shared memory: up to %d pixels (%d for multiplet regions and the "
                "adaptive step method); a region of %d pixels was given.  Fit it as several regions "
                "(VPspectrum.split_long_regions cuts where the flux is closest to the continuum and flags cuts "
                "through absorption), or evaluate likelihoods of long arrays with engine.loglike (K1 takes any length)"
                % (MAX_SAMPLER_PIXELS, TILE_PIXELS, longest))
        self.cmax = int(cmax or self.max_ncomp)
        self.ts = 4 * self.cmax + 1
        self.np_per = 4 if self.voigt else 3
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.burn, self.thin = int(burn), int(thin)
        self.tune_interval, self.tune_throughout = int(tune_interval), int(bool(tune_throughout))
        self.trace_cap = int(trace_cap)
        # "metropolis": PyMC 2's scalar Metropolis per stochastic (what the reference runs); "adaptive": PyMC 2's
        # AdaptiveMetropolis per component (named by mcmc_fit, vpfits.py:361) - include/vamp_b200.h, step_kind
        if step_method not in ("metropolis", "adaptive"):
            raise ValueError("step_method must be 'metropolis' or 'adaptive'")
        self.step_kind = _lib.VAMP_STEP_ADAPTIVE if step_method == "adaptive" else _lib.VAMP_STEP_METROPOLIS
        self.am_delay, self.am_interval = int(am_delay), int(am_interval)
        self.iteration = 0
        dev = self.ctx.torch_device
        B, ts = self.B, self.ts
        f64, f32, i32 = torch.float64, torch.float32, torch.int32
        self.t = {
            "item_region": torch.from_numpy(self.item_region_h).to(dev),
            "item_ncomp": torch.from_numpy(self.item_ncomp_h).to(dev),
            "chain_uid": torch.from_numpy(np.asarray(chain_uid, dtype=np.uint64).view(np.int64)).to(dev),
            "theta": torch.zeros((B, ts), dtype=f64, device=dev),
            "asf": torch.ones((B, ts), dtype=f32, device=dev),
            "psd": torch.ones((B, ts), dtype=f32, device=dev),
            "acc": torch.zeros((B, ts), dtype=i32, device=dev),
            "rej": torch.zeros((B, ts), dtype=i32, device=dev),
            "acc_total": torch.zeros((B, ts), dtype=i32, device=dev),
            "logp": torch.zeros(B, dtype=f64, device=dev),
            "lnL": torch.zeros(B, dtype=f64, device=dev),
            "best_theta": torch.zeros((B, ts), dtype=f64, device=dev),
            "best_logp": torch.zeros(B, dtype=f64, device=dev),
            "mean": torch.zeros((B, ts), dtype=f64, device=dev),
            "m2": torch.zeros((B, ts), dtype=f64, device=dev),
            "comom": torch.zeros((B, self.cmax, 6), dtype=f64, device=dev),
            "ntally": torch.zeros(B, dtype=i32, device=dev),
            "sum_lnL": torch.zeros(B, dtype=f64, device=dev),
            "sum_lnL2": torch.zeros(B, dtype=f64, device=dev),
            "trace": torch.zeros((B, self.trace_cap, ts), dtype=f64, device=dev) if self.trace_cap > 0 else None,
            "work": torch.zeros((B, 3), dtype=torch.int64, device=dev),
            "am_state": (torch.zeros((B, self.cmax, _lib.VAMP_AM_STRIDE), dtype=f64, device=dev)
                         if self.step_kind == _lib.VAMP_STEP_ADAPTIVE else None),
        }

    def _args(self, it0=0, it1=0, rows=None, max_ncomp=None):
        """``rows = (a, b)``: the launch covers sampler rows [a, b) only (every per-chain array is
        addressed from row a) - used to run model groups on separate streams."""
        a0, b0 = (0, self.B) if rows is None else (int(rows[0]), int(rows[1]))
        a = McmcArgs()
        a.n_chains, a.cmax, a.sd_free = b0 - a0, self.cmax, int(self.sd_free)
        a.max_ncomp = self.max_ncomp if max_ncomp is None else int(max_ncomp)
        a.it0, a.it1, a.burn, a.thin = int(it0), int(it1), self.burn, self.thin
        a.tune_interval, a.tune_throughout, a.seed = self.tune_interval, self.tune_throughout, self.seed
        for name in ("item_region", "item_ncomp", "chain_uid", "theta", "asf", "psd", "acc", "rej", "acc_total", "logp",
                     "lnL", "best_theta", "best_logp", "mean", "m2", "comom", "ntally", "sum_lnL", "sum_lnL2"):
            t = self.t[name]
            setattr(a, name, t.data_ptr() + a0 * t.stride(0) * t.element_size())
        tr = self.t["trace"]
        a.trace = (tr.data_ptr() + a0 * tr.stride(0) * tr.element_size()) if tr is not None else None
        a.trace_cap = self.trace_cap
        w = self.t["work"]
        a.work = w.data_ptr() + a0 * w.stride(0) * w.element_size()
        a.step_kind, a.am_delay, a.am_interval = self.step_kind, self.am_delay, self.am_interval
        # a chain whose region has an odd number of 64-pixel row pairs always runs the pair-tail kernel
        pairs = (self.regions.npix[self.item_region_h[a0:b0]] + 63) // 64
        a.pair_tail = int(self.regions.multiplet is None and bool(np.any(pairs & 1)))
        # the launch's longest region decides between the one-tile kernels and the long-region ones
        a.max_npix = int(self.regions.npix[self.item_region_h[a0:b0]].max()) if b0 > a0 else 0
        am = self.t["am_state"]
        a.am_state = (am.data_ptr() + a0 * am.stride(0) * am.element_size()) if am is not None else None
        return a

    def init(self, theta0_scaled=None):
        """PyMC 2-style start (prior draws) or explicit ``theta0_scaled [B, ts]``."""
        torch = _torch()
        self.ctx.use_current_stream()
        if theta0_scaled is not None:
            self.t["theta"].copy_(torch.as_tensor(np.asarray(theta0_scaled, dtype=np.float64)))
        a = self._args()
        check(self.lib.vamp_mcmc_init_dev(self.ctx.handle, self.regions.handle, self.kind, C.byref(a),
                                          0 if theta0_scaled is not None else 1))
        self.iteration = 0

    def run(self, iters, rows=None, max_ncomp=None, advance=True):
        """Advance every chain (or sampler rows ``[a, b)``) by ``iters`` iterations, asynchronously on
        the current torch stream.  ``advance=False`` leaves the iteration counter alone (the caller
        launches several row groups for the same iteration range)."""
        self.ctx.use_current_stream()
        a = self._args(self.iteration, self.iteration + int(iters), rows, max_ncomp)
        check(self.lib.vamp_mcmc_run_dev(self.ctx.handle, self.regions.handle, self.dtype, self.kind, C.byref(a)))
        if advance:
            self.iteration += int(iters)

    def n_free(self, ncomp):
        return self.np_per * ncomp + (1 if self.sd_free else 0)

    def voigt_evals(self, iters):
        """Profile (Voigt/Gauss) pixel-component evaluations *actually computed* by ``iters``
        iterations: every centroid / width step re-evaluates one component over the region;
        amplitude and sd steps re-evaluate none."""
        P = self.regions.npix[self.item_region_h].astype(np.int64)
        return int(np.sum(P * self.item_ncomp_h.astype(np.int64) * (self.np_per - 1)) * int(iters))

    def work(self):
        """Work actually done so far (device counters): dict of pixel-level totals
        ``shape_pix`` (profile pixel-component evaluations), ``far_pix`` (those on the
        continued-fraction path), ``like_pix`` (exp(-tau)+residual pixel passes)."""
        w = self.t["work"].cpu().numpy().astype(np.float64)
        P = self.regions.npix[self.item_region_h].astype(np.float64)
        return {"shape_pix": float(np.sum(w[:, 0] * P)), "far_pix": float(np.sum(w[:, 1])),
                "like_pix": float(np.sum(w[:, 2] * P))}

    def loglike_evals(self, iters):
        """Pixel evaluations of exp(-tau) + residual (one per scalar step per pixel)."""
        P = self.regions.npix[self.item_region_h].astype(np.int64)
        k = self.np_per * self.item_ncomp_h.astype(np.int64) + (0)
        return int(np.sum(P * k) * int(iters))

    def map_polish(self, starts_scaled, item_region, item_ncomp, maxiter=3000, maxfun=0, xtol=1e-4, ftol=1e-3,
                   dtype=None):
        """K3 on explicit starts ``[S, ts]`` (scaled).  Returns dict of numpy arrays."""
        torch = _torch()
        self.ctx.use_current_stream()
        dev = self.ctx.torch_device
        th = torch.as_tensor(np.ascontiguousarray(starts_scaled, dtype=np.float64)).to(dev).contiguous()
        return self._map(th, torch.as_tensor(np.asarray(item_region, dtype=np.int32)).to(dev),
                         torch.as_tensor(np.asarray(item_ncomp, dtype=np.int32)).to(dev),
                         int(np.max(item_ncomp)), maxiter, maxfun, xtol, ftol, dtype)

    def _map(self, th, d_ir, d_in, max_ncomp, maxiter, maxfun, xtol, ftol, dtype=None):
        torch = _torch()
        self.ctx.use_current_stream()
        dev = self.ctx.torch_device
        S = th.shape[0]
        out = {k: torch.empty(S, dtype=torch.float64, device=dev) for k in ("logp", "lnL", "chi2", "bic")}
        out["iters"] = torch.empty(S, dtype=torch.int32, device=dev)
        out["nfev"] = torch.empty(S, dtype=torch.int32, device=dev)
        a = MapArgs()
        a.n_starts, a.cmax, a.max_ncomp, a.sd_free = S, self.cmax, int(max_ncomp), int(self.sd_free)
        a.maxiter, a.maxfun, a.xtol, a.ftol = int(maxiter), int(maxfun), float(xtol), float(ftol)
        a.item_region, a.item_ncomp, a.theta = d_ir.data_ptr(), d_in.data_ptr(), th.data_ptr()
        a.out_logp, a.out_lnL, a.out_chi2, a.out_bic = (out["logp"].data_ptr(), out["lnL"].data_ptr(),
                                                        out["chi2"].data_ptr(), out["bic"].data_ptr())
        a.out_iters, a.out_nfev = out["iters"].data_ptr(), out["nfev"].data_ptr()
        # "mixed": fp32 objective in the simplex search, float64 simplex + float64 report, one warp per start
        code = self.dtype if dtype is None else (VAMP_MIXED if dtype == "mixed" else dtype_code(dtype))
        check(self.lib.vamp_map_dev(self.ctx.handle, self.regions.handle, code, self.kind, C.byref(a)))
        out["theta"] = th
        return out

    def get(self, name):
        return self.t[name].cpu().numpy()
