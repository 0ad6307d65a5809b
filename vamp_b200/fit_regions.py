"""``fit_regions`` - fit many independent absorption regions, sharded over the GPUs of a node.

The reference fits the regions of a spectrum one after the other in a Python loop
(vamp_1.0/vpspectrum.py:273-430) and parallelises only by process-per-spectrum-file
(vamp_1.0/do_vamp.py:85-96).  Regions share nothing, so here they are the shard axis
(SURVEY.md section 8e): one process per GPU, every rank fits its own regions with the batched
kernels (all candidate component counts x restarts x chains of all its regions in one K2 and
one K3 launch per round), and a single fixed-size-record all-gather (NCCL over NVLink when the
process group is NCCL) assembles the per-region BIC tables and posterior summaries on every
rank.  There is no per-iteration communication.  Philox streams are keyed by the *global*
region uid, so a region's result is independent of world size and placement.
"""
import numpy as np

from . import batch, engine


def _gauss3_weights():
    """The kernel ``scipy.ndimage.gaussian_filter(x, 3)`` correlates with (sigma 3, truncate 4 -> radius 12)."""
    x = np.arange(-12, 13)
    phi = np.exp(-0.5 / 9.0 * x ** 2)
    return (phi / phi.sum())[::-1].copy()


_G3 = _gauss3_weights()


def _smooth3(flux_array):
    """``gaussian_filter(flux, 3)`` bit for bit (it calls ``correlate1d`` with these weights, mode 'reflect') without
    the wrappers' argument handling, which costs 10 x the arithmetic on a 100-pixel region."""
    from scipy.ndimage import correlate1d
    return correlate1d(np.asarray(flux_array, dtype=np.float64), _G3, mode="reflect")


def estimate_n(flux_array):
    """Initial component count, as ``VPregion.estimate_n`` (vamp_1.0/vpregion.py:21-35): the local minima of
    ``gaussian_filter(flux, 3)`` (``argrelextrema(.., np.less)``: strictly below both neighbours, never an end
    pixel), 1 when there are fewer than four.  Same numbers as the scipy calls of the reference (tests), at 1/25 of
    their cost - 10^5 regions would otherwise spend 50 s of host time here on every rank."""
    f = _smooth3(flux_array)
    k = int(np.count_nonzero((f[1:-1] < f[:-2]) & (f[1:-1] < f[2:]))) if f.size > 2 else 0
    return k if k >= 4 else 1


def estimate_n_doublet(flux_array, noise_array, n_sigma=3.0):
    """Initial absorber count of a DOUBLET region (new; the reference has single lines only).  A region holding both
    lines of a doublet is mostly continuum, where ``estimate_n``'s count of every local minimum of the smoothed flux
    counts noise; here only minima ``n_sigma`` smoothed-noise deviations below the continuum count, and two of them
    make one absorber."""
    f = _smooth3(flux_array)
    idx = 1 + np.nonzero((f[1:-1] < f[:-2]) & (f[1:-1] < f[2:]))[0] if f.size > 2 else np.zeros(0, dtype=np.int64)
    sm_noise = float(np.median(np.abs(noise_array))) / np.sqrt(2.0 * np.sqrt(np.pi) * 3.0)   # noise after the sigma = 3 filter
    k = int(np.sum(1.0 - f[idx] > n_sigma * sm_noise))
    return max(1, (k + 1) // 2)


def region_cost(npix, n0, window, voigt):
    """Relative cost of fitting a region: pixels x sum over candidates of free parameters."""
    per = 4 if voigt else 3
    return float(npix) * sum(per * n + 1 for n in range(n0, n0 + window))


def region_costs(npix, n0, window, voigt):
    """``region_cost`` of many regions at once (numpy)."""
    per = 4 if voigt else 3
    n0 = np.asarray(n0, dtype=np.float64)
    w = float(window)
    return np.asarray(npix, dtype=np.float64) * (per * (w * n0 + 0.5 * w * (w - 1.0)) + w)


def lpt_partition(costs, world_size):
    """Longest-processing-time-first static partition.  Returns a list of index lists
    (deterministic: ties broken by index).  Heap-based: O(R log W), 10^5 regions in ~0.1 s."""
    import heapq
    costs = np.asarray(costs, dtype=np.float64)
    order = np.lexsort((np.arange(costs.size), -costs))
    heap = [(0.0, k) for k in range(world_size)]
    parts = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        parts[r].append(int(i))
        heapq.heappush(heap, (load + float(costs[i]), r))
    return [sorted(p) for p in parts]


# ``fit_regions(..., **REFERENCE_DEFAULTS)``: the reference's own sampling settings - ONE chain per restart
# (vpfits.py:417-425), PyMC 2's tuning interval of 1000 iterations, Nelder-Mead from the chain's LAST state
# (vpfits.py:426).  The many-chain defaults explore better (lower BIC per candidate); this mode exists so that the
# ladder's outcome can be compared with the reference's like for like (tests/test_ladder_gpu.py).
REFERENCE_DEFAULTS = dict(chains_per_restart=1, tune_interval=1000, polish_from="last", polish_top=1)

MIXED_POLISH_STARTS = 2048
MAX_BATCH_REGIONS = 16384      # regions per device batch: ~1.2e6 chains, ~5 GB of chain state at the C4 settings


def per_region_multiplet(multiplet, R):
    """``multiplet`` (None, one tuple for every region, or a per-region list) -> per-region list or None."""
    if multiplet is None:
        return None
    if len(multiplet) in (2, 3) and np.isscalar(multiplet[0]):
        return [tuple(multiplet)] * R
    if len(multiplet) != R:
        raise ValueError("multiplet: one (freq_ratio, amp_ratio[, lorentz_ratio]) tuple, or one per region")
    return list(multiplet)


def fit_chunked(fit_fn, regions_host, uids, n0, max_batch_regions=MAX_BATCH_REGIONS, **kw):
    """A shard larger than ``max_batch_regions`` is fitted in consecutive device batches (bounded chain state in
    HBM whatever the shard size - 10^5 regions on one GPU are seven batches); Philox streams are keyed by region
    uid, so the batching does not change any result."""
    R = len(regions_host)
    step = max(1, int(max_batch_regions))
    if R <= step:
        return fit_fn(regions_host, uids, n0=n0, **kw)
    mult = kw.pop("multiplet", None)
    out = []
    for a in range(0, R, step):
        if mult is not None:
            kw["multiplet"] = mult[a: a + step]
        out.append(fit_fn(regions_host[a: a + step], uids[a: a + step], n0=n0[a: a + step], **kw))
    return np.concatenate(out, axis=0)


def default_polish_top(chains_per_restart):
    """How many of a restart group's chains get the Nelder-Mead polish (the reference polishes its single chain,
    vpfits.py:419-426): the best quarter, at least 1 and at most 4."""
    return max(1, min(4, int(chains_per_restart) // 4))


class RegionResult:
    """Per-region outcome: chosen ``n``, the BIC / reduced-chi^2 ladder, the MAP parameters of
    the kept fit (physical units) and posterior mean / std of its winning chain."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


_HEAD = 7       # uid, status, n0, kept n, exit n, candidates fitted, tallied samples of the winning chain


def _record_len(nmax, G, cmax):
    return _HEAD + 2 * nmax * G + 1 + 3 * cmax * 4 + 3 + 1 + cmax


def pack_record(uid, status, n0, kept, exit_n, tables, nmax, G, cmax, theta, sd, mean, std, sd_mean, sd_std, lnL,
                chi2=np.nan, cov_aw=None, ntally=0):
    """Fixed-geometry float64 record of one region (the unit of the result all-gather).  ``chi2``:
    noise-weighted chi^2 of the kept fit; ``cov_aw``: posterior covariance of (amplitude, Gaussian
    width) per component (physical units) - what ``chain_covariance`` / ``ErrorN`` consume."""
    rec = np.full(_record_len(nmax, G, cmax), np.nan)
    rec[0:_HEAD] = [uid, status, n0, kept, exit_n, len(tables), ntally]
    bic = np.full((nmax, G), np.nan)
    red = np.full((nmax, G), np.nan)
    for n, (b, r) in tables.items():
        if n - n0 < nmax:
            bic[n - n0], red[n - n0] = b, r
    o = _HEAD
    rec[o: o + nmax * G] = bic.reshape(-1); o += nmax * G
    rec[o: o + nmax * G] = red.reshape(-1); o += nmax * G
    rec[o] = lnL; o += 1
    for arr in (theta, mean, std):
        a = np.full((cmax, 4), np.nan)
        a[: arr.shape[0]] = arr[:cmax]
        rec[o: o + cmax * 4] = a.reshape(-1); o += cmax * 4
    rec[o: o + 3] = [np.nan if sd is None else sd, np.nan if sd_mean is None else sd_mean,
                     np.nan if sd_std is None else sd_std]
    o += 3
    rec[o] = chi2
    if cov_aw is not None:
        k = min(len(cov_aw), cmax)
        rec[o + 1: o + 1 + k] = np.asarray(cov_aw, dtype=np.float64)[:k]
    return rec


def unpack_record(rec, nmax, G, cmax):
    uid, status, n0, kept, exit_n, ncand, ntally = (int(v) for v in rec[0:_HEAD])
    o = _HEAD
    bic = rec[o: o + nmax * G].reshape(nmax, G); o += nmax * G
    red = rec[o: o + nmax * G].reshape(nmax, G); o += nmax * G
    lnL = rec[o]; o += 1
    arrs = []
    for _ in range(3):
        arrs.append(rec[o: o + cmax * 4].reshape(cmax, 4)[:kept].copy()); o += cmax * 4
    sd, sd_mean, sd_std = rec[o: o + 3]
    o += 3
    chi2 = rec[o]
    cov_aw = rec[o + 1: o + 1 + cmax][:kept].copy()
    return RegionResult(uid=uid, status=status, n0=n0, n=kept, exit_n=exit_n, chi2=chi2, cov_aw=cov_aw, ntally=ntally,
                        bic={n0 + i: bic[i].copy() for i in range(min(ncand, nmax))},
                        red_chi={n0 + i: red[i].copy() for i in range(min(ncand, nmax))},
                        lnL=lnL, theta=arrs[0], mean=arrs[1], std=arrs[2], sd=sd, sd_mean=sd_mean, sd_std=sd_std)


# ladder rounds with fewer models than this fit one candidate AHEAD of the rule (see fit_local); 256 models x 24 chains
# are two waves of sampler CTAs on a B200
SPEC_BELOW_MODELS = int(__import__("os").environ.get("VAMP_SPEC_BELOW", "256"))


def fit_local(regions_host, uids, voigt=False, sd_free=True, iterations=3000, burn=300, thin=15, restarts=3,
              chains_per_restart=32, window=3, max_rounds=None, chi_limit=1.5, dtype="f32", seed=0, lsf=None,
              n0=None, nmax=12, ctx=None, polish_top=None, tune_interval=50, stats=None, cmax_rec=None, polish_dtype="f64",
              multiplet=None, polish_from="best", step_method="metropolis", spec_below=None):
    """Fit the given regions on this process's GPU.  Returns ``[R, record_len]`` float64.
    ``multiplet``: per-region ``(freq_ratio, amp_ratio[, lorentz_ratio])`` or None (``engine.Regions``);
    ``polish_from``: see ``batch.fit_models``.
    ``window``: candidate counts fitted per ladder round for every region still undecided - an int, or
    ``(first, later)`` (e.g. ``(2, 1)``: the two candidates the first comparison needs, then one more per round: no
    speculative fits at all, as the reference's sequential ladder, at the price of more rounds).  The ladder's
    decisions do not depend on it (chains are keyed by (region, n)); ``max_rounds`` defaults to what ``nmax`` needs.
    ``spec_below``: a late round whose models would not fill the GPU anyway (fewer than this many; default
    ``SPEC_BELOW_MODELS``, 0 = never) fits one candidate more per region than the rule needs next - such a round lasts
    as long as its longest chain whatever it holds, so the look-ahead is free and halves the number of these rounds.
    Candidates the rule never looks at are left out of the record (``batch.ladder_examined``)."""
    R = len(regions_host)
    G = int(restarts)
    if R == 0:                                   # nothing to fit (an empty shard): no device work, no context
        cm = int(cmax_rec or 1 + nmax)
        return np.zeros((0, _record_len(nmax, G, cm)))
    ctx = ctx or engine.default_context()
    if polish_top is None:                       # chains polished per restart group: a quarter of them, 1 .. 4
        polish_top = default_polish_top(chains_per_restart)
    n0 = [estimate_n(r[1]) for r in regions_host] if n0 is None else [int(v) for v in n0]
    cap = [max(1, (len(r[1]) - 1) // 2) for r in regions_host]
    freedom = [len(regions_host[i][1]) - 3 * n0[i] for i in range(R)]        # vpregion.py:37-39, frozen at n0
    cmax_rec = int(cmax_rec or (max(n0) + nmax if R else 1 + nmax))
    if R == 0:
        return np.zeros((0, _record_len(nmax, G, cmax_rec)))
    import time as _time
    _t = _time.perf_counter()
    regs = engine.Regions(ctx, regions_host, lsf=lsf, multiplet=multiplet)
    if stats is not None:
        stats["h2d_bytes"] = stats.get("h2d_bytes", 0) + regs.h2d_bytes
        tm = stats.setdefault("timing", {})
        tm["upload_s"] = tm.get("upload_s", 0.0) + (_time.perf_counter() - _t)
    tables = [dict() for _ in range(R)]
    fits = [dict() for _ in range(R)]
    nxt = list(n0)
    active = list(range(R))
    outcome = [None] * R
    w_first, w_later = (int(window), int(window)) if np.isscalar(window) else (int(window[0]), int(window[1]))
    if max_rounds is None:
        max_rounds = 1 + max(0, -(-(nmax - w_first) // max(w_later, 1)))
    for rnd in range(max_rounds):
        if not active:
            break
        models, owner = [], []
        wnd = w_first if rnd == 0 else w_later
        if rnd > 0 and len(active) * wnd < (SPEC_BELOW_MODELS if spec_below is None else int(spec_below)):
            wnd += 1
        for i in active:
            for n in range(nxt[i], min(nxt[i] + wnd, n0[i] + nmax)):
                if n <= cap[i] or n == nxt[i]:
                    models.append((i, n))
                    owner.append(i)
        mb = batch.fit_models(regs, models, voigt=voigt, sd_free=sd_free, iterations=iterations, burn=burn, thin=thin,
                              restarts=G, chains_per_restart=chains_per_restart, polish_top=polish_top, dtype=dtype,
                              seed=seed, region_uids=uids, tune_interval=tune_interval, polish_dtype=polish_dtype,
                              polish_from=polish_from, step_method=step_method,
                              timing=None if stats is None else stats.setdefault("timing", {}))
        if stats is not None and stats.get("trace_rounds"):
            tmr = stats["timing"]
            stats.setdefault("rounds", []).append((rnd, len(active), len(models), max(n for _, n in models),
                                                   round(tmr.get("k2_ms", 0.0)), round(tmr.get("polish_ms", 0.0)),
                                                   round(_time.perf_counter() - stats.setdefault("t0", _time.perf_counter()), 3)))
        if stats is not None:
            stats["voigt_evals"] = stats.get("voigt_evals", 0) + mb.voigt_evals
            stats["voigt_evals_actual"] = stats.get("voigt_evals_actual", 0) + mb.work["shape_pix"]
            stats["launches"] = stats.get("launches", 0) + 4
            stats["chains"] = stats.get("chains", 0) + mb.n_chains
        _t = _time.perf_counter()
        for m, (i, n) in enumerate(models):
            tables[i][n] = (mb.bic[m].copy(), mb.chi2[m] / freedom[i])
            fits[i][n] = (mb, m)
        still = []
        for i in active:
            kept, exit_n, decided = batch.ladder_replay(n0[i], tables[i], chi_limit)
            last = max(tables[i])
            if decided or last >= min(cap[i], n0[i] + nmax - 1):
                outcome[i] = (kept, exit_n if decided else kept, 0 if decided else 1)
            else:
                nxt[i] = last + 1
                still.append(i)
        active = still
        if stats is not None:
            tm["ladder_s"] = tm.get("ladder_s", 0.0) + (_time.perf_counter() - _t)
    _t = _time.perf_counter()
    for i in active:      # rounds exhausted: keep the best so far
        kept, exit_n, _ = batch.ladder_replay(n0[i], tables[i], chi_limit)
        outcome[i] = (kept, kept, 2)
    out = np.empty((R, _record_len(nmax, G, cmax_rec)))
    for i in range(R):
        kept, exit_n, status = outcome[i]
        if status == 0:                    # decided: drop the candidates fitted ahead of the rule
            last = batch.ladder_examined(n0[i], tables[i], chi_limit)
            tables[i] = {n: v for n, v in tables[i].items() if n <= last}
        mb, m = fits[i][kept]
        g = int(np.nanargmin(mb.bic[m])) if np.any(np.isfinite(mb.bic[m])) else 0
        # covariance of (amplitude, Gaussian width) from the co-moments of the winning chain: Gauss (A, sigma) =
        # pair 1, Voigt (A, G_fwhm) = pair 2 converted to sigma; widths are in scaled units on the device
        pair = 2 if voigt else 1
        fac = mb.scale[m] * (1.0 / (2.0 * np.sqrt(2.0 * np.log(2.0))) if voigt else 1.0)
        cov_aw = mb.comom[m, g, :kept, pair] / max(int(mb.ntally[m, g]) - 1, 1) * fac      # np.cov normalisation
        out[i] = pack_record(uids[i], status, n0[i], kept, exit_n, tables[i], nmax, G, cmax_rec,
                             mb.theta[m, g, :kept], None if mb.sd is None else mb.sd[m, g], mb.mean[m, g, :kept],
                             mb.std[m, g, :kept], None if mb.sd_mean is None else mb.sd_mean[m, g],
                             None if mb.sd_std is None else mb.sd_std[m, g], mb.lnL[m, g], chi2=mb.chi2[m, g],
                             cov_aw=cov_aw, ntally=int(mb.ntally[m, g]))
    _t2 = _time.perf_counter()
    regs.close()                                  # the next device batch (fit_chunked) starts from free memory
    if stats is not None:
        tm["records_s"] = tm.get("records_s", 0.0) + (_t2 - _t)
        tm["close_s"] = tm.get("close_s", 0.0) + (_time.perf_counter() - _t2)
    return out


def gather_records(local, world_size, rec_len, device=None):
    """All-gather of the ranks' ``[R_r, rec_len]`` record blocks (padded to the longest) with
    ``torch.distributed`` - NCCL over NVLink on GPUs, gloo in the CPU tests."""
    import torch
    import torch.distributed as dist
    dev = device or "cpu"
    counts = torch.zeros(world_size, dtype=torch.int64, device=dev)
    mine = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine)
    rmax = int(counts.max().item())
    buf = torch.full((rmax, rec_len), float("nan"), dtype=torch.float64, device=dev)
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    allbuf = torch.empty((world_size * rmax, rec_len), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(allbuf, buf)
    allbuf = allbuf.cpu().numpy().reshape(world_size, rmax, rec_len)
    cnt = counts.cpu().numpy()
    return np.concatenate([allbuf[r, : cnt[r]] for r in range(world_size)], axis=0)


class RegionResults:
    """The uid-ordered result of ``fit_regions``: a read-only sequence of :class:`RegionResult` that unpacks a region's
    fixed-geometry record on access (10^5 regions cost nothing until they are looked at; every rank of a
    distributed fit holds the same records).  ``records`` is the ``[R, record_len]`` float64 array itself."""

    def __init__(self, records, nmax, G, cmax):
        self.records, self.nmax, self.G, self.cmax = records, int(nmax), int(G), int(cmax)

    def __len__(self):
        return int(self.records.shape[0])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        return unpack_record(self.records[i], self.nmax, self.G, self.cmax)

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def __eq__(self, other):
        return list(self) == other if isinstance(other, list) else NotImplemented

    @property
    def n(self):
        """Chosen component count of every region (int array) without unpacking anything."""
        return self.records[:, 3].astype(np.int64)

    @property
    def uid(self):
        return self.records[:, 0].astype(np.int64)

    @property
    def status(self):
        return self.records[:, 1].astype(np.int64)

    def checksum(self):
        """SHA-256 of the records (NaN padding included: it is produced by the same code on every rank).  Philox
        streams are keyed by region uid and every kernel schedule is chosen from GLOBAL quantities, so this value is
        the same for every world size and batching of the same call."""
        import hashlib
        return hashlib.sha256(np.ascontiguousarray(self.records, dtype=np.float64).tobytes()).hexdigest()


def _gather_int(local, counts_total, world, device):
    """All-gather of per-rank int64 blocks of known (rank-dependent) lengths ``counts_total[rank]``."""
    import torch
    import torch.distributed as dist
    dev = device or "cpu"
    m = int(max(counts_total)) if len(counts_total) else 0
    buf = torch.zeros(max(m, 1), dtype=torch.int64, device=dev)
    if len(local):
        buf[: len(local)] = torch.as_tensor(np.asarray(local, dtype=np.int64), device=dev)
    allb = torch.empty(world * max(m, 1), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allb, buf)
    allb = allb.cpu().numpy().reshape(world, max(m, 1))
    return np.concatenate([allb[r, : counts_total[r]] for r in range(world)]) if world else np.zeros(0, np.int64)


def initial_counts(regions_host, multiplet, lo=0, hi=None):
    """``estimate_n`` (doublet regions: ``estimate_n_doublet``) of regions ``[lo, hi)``."""
    hi = len(regions_host) if hi is None else hi
    out = []
    for i in range(lo, hi):
        m = None if multiplet is None else multiplet[i]
        if m is not None and m[1] != 0.0:           # a doublet shows two minima per absorber
            out.append(estimate_n_doublet(regions_host[i][1], regions_host[i][2]))
        else:
            out.append(estimate_n(regions_host[i][1]))
    return out


def shard_plan(npix, n0, window, voigt, world):
    """The static partition of the regions over ``world`` ranks (LPT on ``region_costs``): list of index lists."""
    return lpt_partition(region_costs(npix, n0, window, voigt), world)


def fit_regions(regions_host, uids=None, distributed=None, fit_fn=None, emulate=None, **kw):
    """Fit ``regions_host = [(nu, flux, noise), ...]`` (host float64 arrays, nu ascending).

    With an initialised ``torch.distributed`` process group (or ``distributed=True``) every rank estimates the
    initial component count of a 1/W slice of the regions (one small all-gather makes the counts global), the
    regions are LPT-partitioned over the ranks, each rank fits its share on its own GPU and the fixed-size records
    are all-gathered; every rank returns the full, uid-ordered :class:`RegionResults`.
    ``fit_fn`` (tests only) replaces the GPU fit of a shard.  ``emulate=(rank, world)`` (tests): no process group -
    compute exactly the records rank ``rank`` of a ``world``-rank run would compute, and return only those."""
    import torch.distributed as dist
    R = len(regions_host)
    uids = list(range(R)) if uids is None else [int(u) for u in uids]
    voigt = kw.get("voigt", False)
    window = kw.get("window", 3)
    window = int(window) if np.isscalar(window) else int(window[0])      # first-round window: cost model, K3 schedule
    nmax = kw.get("nmax", 12)
    G = kw.get("restarts", 3)
    n0 = kw.pop("n0", None)
    use_dist = dist.is_available() and dist.is_initialized() if distributed is None else bool(distributed)
    if emulate is not None:
        use_dist = False
    mult = per_region_multiplet(kw["multiplet"], R) if kw.get("multiplet") is not None else None
    device = None
    if use_dist:
        world, rank = dist.get_world_size(), dist.get_rank()
        if dist.get_backend() == "nccl":
            import torch
            device = torch.device("cuda", torch.cuda.current_device())
    if n0 is None:
        if use_dist and R >= 4 * world:
            # every rank counts the minima of its slice only; the counts travel, not the work
            bounds = [(r * R) // world for r in range(world + 1)]
            mine = initial_counts(regions_host, mult, bounds[rank], bounds[rank + 1])
            n0 = [int(v) for v in _gather_int(mine, [bounds[r + 1] - bounds[r] for r in range(world)], world, device)]
        else:
            n0 = initial_counts(regions_host, mult)
    else:
        n0 = [int(v) for v in n0]
    cmax_rec = (max(n0) if n0 else 1) + nmax
    rec_len = _record_len(nmax, G, cmax_rec)
    if kw.get("polish_dtype", "auto") == "auto" and fit_fn is None:
        # K3 schedule from the GLOBAL number of starts (the same on every rank and for every world size):
        # thousands of starts -> one warp per start, fp32 search stage + float64 finish ("mixed");
        # a few -> one CTA per start in float64 (latency)
        starts = R * window * G * (kw.get("polish_top") or default_polish_top(kw.get("chains_per_restart", 32)))
        kw["polish_dtype"] = "mixed" if (starts >= MIXED_POLISH_STARTS and kw.get("lsf") is None) else "f64"
    fit_fn = fit_fn or fit_local
    chunk = kw.pop("max_batch_regions", MAX_BATCH_REGIONS)
    if mult is not None:
        kw["multiplet"] = mult
    else:
        kw.pop("multiplet", None)
    if not use_dist and emulate is None:
        recs = fit_chunked(fit_fn, regions_host, uids, n0, chunk, cmax_rec=cmax_rec, **kw)
    else:
        if emulate is not None:
            rank, world = int(emulate[0]), int(emulate[1])
        npix = np.fromiter((len(r[1]) for r in regions_host), dtype=np.int64, count=R)
        mine = shard_plan(npix, n0, window, voigt, world)[rank]
        if "multiplet" in kw:
            kw["multiplet"] = [kw["multiplet"][i] for i in mine]
        local = fit_chunked(fit_fn, [regions_host[i] for i in mine], [uids[i] for i in mine], [n0[i] for i in mine],
                            chunk, cmax_rec=cmax_rec, **kw)        # same record geometry on every rank
        assert local.shape[1] == rec_len
        recs = local if emulate is not None else gather_records(local, world, rec_len, device)
    order = np.argsort(recs[:, 0].astype(np.int64), kind="stable") if len(recs) else np.zeros(0, np.int64)
    return RegionResults(np.ascontiguousarray(recs[order]) if len(recs) else np.zeros((0, rec_len)), nmax, G, cmax_rec)
