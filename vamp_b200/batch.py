"""Batched model fitting: every (region, candidate component count, restart, chain) at once.

This is the GPU-shaped core under ``VPfit.find_bic``, ``VPregion.region_fit`` and
``fit_regions``: where the reference runs ``3 restarts x (MCMC, MAP)`` sequentially for one
candidate ``n`` at a time (vamp_1.0/vpfits.py:417-428, vpregion.py:58-60), all candidates of all
regions advance inside one K2 launch, their best chains are polished in one K3 launch, and
only the small BIC / chi^2 / MAP tables travel back to the host, where the reference's
acceptance rule is replayed (``ladder_replay``).

Chains are keyed by ``(region uid, n, restart, chain)`` for the Philox stream, so results do
not depend on batching or on how regions are sharded over GPUs.
"""
import numpy as np

from . import engine


def chain_uids(region_uid, n, restarts, chains_per_restart):
    """64-bit Philox chain ids: region uid (40 bits) | n (8 bits) | chain-in-model (16 bits)."""
    c = np.arange(restarts * chains_per_restart, dtype=np.uint64)
    return (np.uint64(int(region_uid) & 0xFFFFFFFFFF) << np.uint64(24)) | (np.uint64(int(n) & 0xFF) << np.uint64(16)) | c


# component-count edges of the polish launch groups: see ``fit_models.polish``
# (None: <= 2, 3..4, then one group per component count - 12 500 C4 regions, polish 3.31 s with the edges (4, 8, 12) -> 3.06 s)
POLISH_GROUP_EDGES = (tuple(int(v) for v in __import__("os").environ["VAMP_POLISH_EDGES"].split(","))
                      if __import__("os").environ.get("VAMP_POLISH_EDGES") else None)


def polish_group_edges(n_largest):
    """Upper component counts of the polish launch groups of a batch whose largest count is ``n_largest``."""
    if POLISH_GROUP_EDGES is not None:
        return tuple(POLISH_GROUP_EDGES) + (int(n_largest),)
    return (2, 4) + tuple(range(5, max(5, int(n_largest)) + 1))
POLISH_STREAMS = __import__("os").environ.get("VAMP_POLISH_STREAMS", "1") != "0"
# Sampler launch groups by component count: K2's shared memory per chain is (n_max + 2) KB with n_max the LAUNCH's largest
# count, and that footprint decides how many CTAs share an SM (<= 6 components: 5, 7: 5, 8..10: 4, 11..14: 3, more: 2).
# Chains of <= 6 components go out together, every larger count as its own launch (VAMP_K2_EDGES: explicit edges;
# VAMP_K2_GROUPS=0: one launch).  12 500 C4 regions, K2: 13.64 s (groups <= 8 / 9-10 / 11-15 / more) -> 13.16 (<= 6 split
# off) -> 12.52 s (one group per count), same records.
K2_GROUP_FIRST = 6
_k2_edges_env = __import__("os").environ.get("VAMP_K2_EDGES")


def k2_group_edges(n_largest):
    """Upper component counts of the sampler launch groups of a batch whose largest count is ``n_largest``."""
    if _k2_edges_env:
        return tuple(int(v) for v in _k2_edges_env.split(",")) + (int(n_largest),)
    return tuple(range(K2_GROUP_FIRST, max(K2_GROUP_FIRST, int(n_largest)) + 1))


K2_GROUPS = __import__("os").environ.get("VAMP_K2_GROUPS", "1") != "0"


_streams = {}


def _polish_streams(dev, n):
    """``n`` side streams of the device (created once)."""
    torch = engine._torch()
    pool = _streams.setdefault(str(dev), [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(dev))
    return pool[:n]


def model_chain_uids(region_uids, models, restarts, chains_per_restart):
    """``chain_uids`` of every model ``(region index, n)``, concatenated in model order, without the per-model loop
    (0.3 s of host time at 4 x 10^4 models)."""
    m_reg = np.asarray(region_uids)[[r for r, _ in models]].astype(np.int64).astype(np.uint64) & np.uint64(0xFFFFFFFFFF)
    m_n = np.array([n for _, n in models], dtype=np.uint64) & np.uint64(0xFF)
    c = np.arange(int(restarts) * int(chains_per_restart), dtype=np.uint64)
    return (((m_reg << np.uint64(24)) | (m_n << np.uint64(16)))[:, None] | c[None, :]).reshape(-1)


class ModelBatch:
    """Result tables of one batched fit (numpy, host).  Shapes: M models, G restarts."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def red_chi(self, freedom):
        """``ReducedChisquared`` per restart (vpfits.py:428) given per-model d.o.f."""
        return self.chi2 / np.asarray(freedom, dtype=np.float64)[:, None]


def fit_models(regions, models, voigt=False, sd_free=True, iterations=3000, burn=300, thin=15, restarts=3,
               chains_per_restart=32, polish_top=4, dtype="f32", seed=0, region_uids=None, tune_interval=50,
               tune_throughout=True, keep_sampler=False, trace=False, maxiter=None, xtol=1e-4, ftol=1e-3,
               polish_dtype="f64", overlap=False, timing=None, polish_from="best", step_method="metropolis"):
    """Fit ``models = [(region_index, n), ...]`` on the device-resident ``regions``.

    Per model and restart group: prior-drawn starts -> ``iterations`` of K2 Metropolis ->
    Nelder-Mead polish (K3) of the group's ``polish_top`` best chains -> the group's MAP,
    ``BIC = k ln P - 2 lnL`` and noise-weighted chi^2 (the quantities ``find_bic`` appends,
    vpfits.py:427-428).  ``polish_from``: "best" - the polish starts from each chain's best visited state (default);
    "last" - from its LAST state, which is what the reference does (``map.fit`` after ``mcmc.sample`` starts from the
    sampler's final values, vpfits.py:425-426).  Returns a :class:`ModelBatch`."""
    import time as _time
    torch = engine._torch()
    _t0 = _time.perf_counter()
    models = [(int(r), int(n)) for r, n in models]
    M, G, Cn = len(models), int(restarts), int(chains_per_restart)
    if M == 0:
        raise ValueError("no models to fit")
    region_uids = np.arange(regions.R, dtype=np.int64) if region_uids is None else np.asarray(region_uids)
    cmax = max(n for _, n in models)
    per = G * Cn
    item_region = np.repeat(np.array([r for r, _ in models], dtype=np.int32), per)
    item_ncomp = np.repeat(np.array([n for _, n in models], dtype=np.int32), per)
    uid = model_chain_uids(region_uids, models, G, Cn)
    # sampler rows sorted by (long class, component count, pixel class, model): CTAs are homogeneous, a kernel launches
    # its rows from the end (most expensive chains first), and the chains of one component count are contiguous,
    # so every count can run as its own launch on its own stream (below).  Long class: regions above one coordinate
    # tile (256 pixels) run the long-region kernels, one launch per buffer length - all 0 for the BASELINE configs.
    ppl = (regions.npix[item_region] + 31) // 32
    lcls = engine.long_class(regions.npix[item_region])
    order = np.lexsort((np.arange(M * per), ppl, item_ncomp, lcls))
    inv = np.empty_like(order)
    inv[order] = np.arange(order.size)
    cap = max(0, (int(iterations) - int(burn) + int(thin) - 1) // max(int(thin), 1) + 1) if trace else 0
    s = engine.Sampler(regions, item_region[order], item_ncomp[order], uid[order], voigt=voigt, sd_free=sd_free,
                       dtype=dtype, seed=seed, burn=burn, thin=thin, tune_interval=tune_interval,
                       tune_throughout=tune_throughout, trace_cap=cap, cmax=cmax, step_method=step_method)
    s.init()
    dev = s.ctx.torch_device
    d_inv = torch.from_numpy(inv).to(dev)
    K = max(1, min(int(polish_top), Cn))
    MG = M * G
    if polish_from not in ("best", "last"):
        raise ValueError("polish_from must be 'best' or 'last'")
    rank_key, start_key = ("best_logp", "best_theta") if polish_from == "best" else ("logp", "theta")
    f64 = torch.float64
    res = {k: torch.empty(MG, dtype=f64, device=dev) for k in ("logp", "lnL", "chi2", "bic")}
    res["iters"] = torch.empty(MG, dtype=torch.int32, device=dev)
    res["theta"] = torch.empty((MG, s.ts), dtype=f64, device=dev)
    primary_rows = torch.empty(MG, dtype=torch.int64, device=dev)

    def polish(model_ids):
        """Top-K chains of every (model, restart) by best log-posterior -> K3 -> winner per (model, restart);
        everything on the device and on the current stream."""
        mid = torch.as_tensor(np.asarray(model_ids, dtype=np.int64), device=dev)
        mg = (mid[:, None] * G + torch.arange(G, device=dev)[None, :]).reshape(-1)        # (model, restart) ids
        chains = (mg[:, None] * Cn + torch.arange(Cn, device=dev)[None, :])               # logical chain ids [mg, Cn]
        rows = d_inv[chains]                                                             # sampler rows
        blp = torch.nan_to_num(s.t[rank_key][rows], nan=-float("inf"))
        top = blp.topk(K, dim=1).indices
        phys = torch.gather(rows, 1, top).reshape(-1)
        # launch order of the polish: cheapest starts first in memory, the kernels take their work from the END of
        # the list (as K2 does), so the long simplexes (many components, 3000 iterations) start first and the tail
        # of the launch is made of short ones
        ncs = s.t["item_ncomp"][phys]
        lpt = torch.argsort(ncs.to(torch.int64) * 1024 + s.t["item_region"][phys].to(torch.int64) % 1024, stable=True)
        unsort = torch.empty_like(lpt)
        unsort[lpt] = torch.arange(lpt.numel(), device=dev)
        ps = phys[lpt]
        starts = s.t[start_key][ps].contiguous()
        p_reg, p_nc = s.t["item_region"][ps].contiguous(), s.t["item_ncomp"][ps].contiguous()
        # The warp-per-start schedule keeps a (k + 1) x (4 n_max + 1) simplex per warp in shared memory, n_max being
        # the LAUNCH's largest component count, and that footprint decides how many warps an SM holds: the starts
        # (sorted by component count) go out in groups of similar size, largest first, so that the many small
        # simplexes do not pay for the few large ones.  Group sizes follow from the model list (no device read).
        ncount = np.bincount(np.array([models[m][1] for m in model_ids]), minlength=2) * (G * K)
        groups, a = [], 0
        if polish_dtype == "mixed":
            for hi in polish_group_edges(len(ncount) - 1):
                b = int(ncount[: hi + 1].sum())
                if b > a:
                    groups.append((a, b, int(np.nonzero(ncount[: hi + 1])[0].max())))
                a = max(a, b)
        else:
            groups.append((0, int(ncount.sum()), int(max(models[m][1] for m in model_ids))))
        parts = []
        if len(groups) == 1:
            a, b, nmax_g = groups[0]
            parts.append(s._map(starts, p_reg, p_nc, nmax_g, int(maxiter or iterations), 0, xtol, ftol, polish_dtype))
        elif not POLISH_STREAMS:
            for a, b, nmax_g in reversed(groups):
                parts.append(s._map(starts[a:b], p_reg[a:b], p_nc[a:b], nmax_g, int(maxiter or iterations), 0, xtol, ftol,
                                    polish_dtype))
            parts.reverse()
        else:
            # one stream per group, largest simplexes first: the groups' tails (a launch lasts as long as its longest
            # simplex) overlap instead of adding up
            cur = torch.cuda.current_stream(dev)
            side = _polish_streams(dev, len(groups))
            for st, (a, b, nmax_g) in zip(side, reversed(groups)):
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    q = s._map(starts[a:b], p_reg[a:b], p_nc[a:b], nmax_g, int(maxiter or iterations), 0, xtol, ftol,
                               polish_dtype)
                for v in q.values():
                    v.record_stream(cur)             # allocated on the side stream, consumed on the current one
                parts.append(q)
            for st in side[: len(groups)]:
                cur.wait_stream(st)
            s.ctx.use_current_stream()
            parts.reverse()
        r = {k: torch.cat([q[k] for q in parts]) for k in parts[0]} if len(parts) > 1 else parts[0]
        r = {k: v[unsort] for k, v in r.items()}
        logp = torch.nan_to_num(r["logp"].view(-1, K), nan=-float("inf"))
        win = logp.argmax(dim=1)
        pick = torch.arange(mg.numel(), device=dev) * K + win
        for k in ("logp", "lnL", "chi2", "bic", "iters", "theta"):
            res[k][mg] = r[k][pick]
        primary_rows[mg] = phys[pick]

    # ---- default: ONE sampler launch over all chains (longest first), then one polish launch.
    #      ``overlap=True`` (experiment, kept for the record): one launch group per component count on its own
    #      stream with the group's polish on a high-priority stream, so that a finished group's K3 (few CTAs,
    #      latency bound) hides under the groups still sampling.  Measured on B200 it is SLOWER - C2 step 472 ->
    #      601 ms (706 ms with stream priorities standing in for the launch order): concurrent grids with
    #      different shared-memory footprints do not share an SM well and the longest-first order is lost.
    counts = sorted({n for _, n in models}, reverse=True)
    sorted_nc = item_ncomp[order]
    main = torch.cuda.current_stream(dev)
    if overlap and len(counts) > 1 and int(lcls.max()) == 0:
        side = []
        for rank, n in enumerate(counts):
            a, b = int(np.searchsorted(sorted_nc, n, "left")), int(np.searchsorted(sorted_nc, n, "right"))
            # stream priorities stand in for the launch order of a single grid: pending CTAs of the more
            # expensive groups are scheduled first (CUDA clamps to the device's priority range, 0 .. -5 on B200)
            st, hp = torch.cuda.Stream(dev, priority=min(0, rank - 4)), torch.cuda.Stream(dev, priority=-5)
            st.wait_stream(main)
            with torch.cuda.stream(st):
                s.run(iterations, rows=(a, b), max_ncomp=n, advance=False)
            hp.wait_stream(st)
            with torch.cuda.stream(hp):
                polish([m for m in range(M) if models[m][1] == n])
            side.append(hp)
        for hp in side:
            main.wait_stream(hp)
        s.iteration += int(iterations)
    else:
        _t1 = _time.perf_counter()
        if timing is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        # K2's shared memory per chain grows with the LAUNCH's largest component count (base + n_max + 1 cache slots of
        # 1 KB): a batch holding a few 17-component chains would run everybody at 2 CTAs per SM instead of 5.  The
        # rows are sorted by component count, so they go out in groups (<= 8, 9..10, 11..15, more), largest first on side
        # streams; a batch whose counts fit one group (BASELINE C2: n <= 6) is the single launch it always was.
        bounds = []
        sorted_lc = lcls[order]
        for lc in np.unique(sorted_lc):                  # one pass per long class (a single pass of class 0 as a rule)
            lo, hi_ = int(np.searchsorted(sorted_lc, lc, "left")), int(np.searchsorted(sorted_lc, lc, "right"))
            nc, a0 = sorted_nc[lo:hi_], 0
            for hi in k2_group_edges(nc[-1]):
                b0 = int(np.searchsorted(nc, hi, "right"))
                if b0 > a0:
                    bounds.append((lo + a0, lo + b0, int(nc[b0 - 1])))
                a0 = max(a0, b0)
        if len(bounds) == 1 or (not K2_GROUPS and sorted_lc[0] == sorted_lc[-1]):
            s.run(iterations)
        elif __import__("os").environ.get("VAMP_K2_GROUP_TIMES"):
            # (diagnostic: the groups one after the other, each timed - how the sampler's time splits over component counts)
            for a0, b0, nmax_g in reversed(bounds):
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                s.run(iterations, rows=(a0, b0), max_ncomp=nmax_g, advance=False)
                g1.record()
                torch.cuda.synchronize()
                px = regions.npix[item_region[order][a0:b0]]
                print("   K2 group n <= %2d: %7d chains (%.0f%% of <= 128 px) %8.1f ms" % (
                    nmax_g, b0 - a0, 100.0 * float(np.mean(px <= 128)), g0.elapsed_time(g1)), flush=True)
            s.iteration += int(iterations)
        else:
            side = _polish_streams(dev, len(bounds))
            for st, (a0, b0, nmax_g) in zip(side, reversed(bounds)):
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    s.run(iterations, rows=(a0, b0), max_ncomp=nmax_g, advance=False)
            for st in side:
                main.wait_stream(st)
            s.ctx.use_current_stream()
            s.iteration += int(iterations)
        if timing is not None:
            ev[1].record()
        polish(list(range(M)))
        if timing is not None:
            ev[2].record()
    _t2 = _time.perf_counter()
    out = {k: res[k].view(M, G).cpu().numpy() for k in ("logp", "lnL", "chi2", "bic", "iters")}
    _t3 = _time.perf_counter()
    if timing is not None and not (overlap and len(counts) > 1):
        # host seconds: set-up before the launches, launching, waiting for the device; device ms of K2 and the polish
        for key, v in (("host_setup_s", _t1 - _t0), ("host_launch_s", _t2 - _t1), ("host_wait_s", _t3 - _t2),
                       ("k2_ms", ev[0].elapsed_time(ev[1])), ("polish_ms", ev[1].elapsed_time(ev[2]))):
            timing[key] = timing.get(key, 0.0) + v
    theta_map = res["theta"].cpu().numpy()                                    # [M*G, ts] scaled
    prow = primary_rows.cpu().numpy()                                         # sampler row of the winning chain
    ts = s.ts
    comps_s = theta_map[:, : 4 * cmax].reshape(M * G, cmax, 4)
    reg_of = np.repeat(np.array([r for r, _ in models]), G)
    comps = regions.to_physical(comps_s, reg_of).reshape(M, G, cmax, 4)
    sd = theta_map[:, 4 * cmax].reshape(M, G) if sd_free else None
    # posterior summaries of the winning chains (Welford accumulators of K2)
    ntal = s.t["ntally"][primary_rows].cpu().numpy().reshape(M, G)
    mean_s = s.t["mean"][primary_rows].cpu().numpy()
    m2_s = s.t["m2"][primary_rows].cpu().numpy()
    com = s.t["comom"][primary_rows].cpu().numpy().reshape(M, G, cmax, 6)
    sc = regions.scale[reg_of][:, None]
    mean_c = regions.to_physical(mean_s[:, : 4 * cmax].reshape(M * G, cmax, 4), reg_of).reshape(M, G, cmax, 4)
    with np.errstate(invalid="ignore", divide="ignore"):
        var_s = m2_s / np.maximum(ntal.reshape(-1, 1), 1)
    std_c = np.sqrt(var_s[:, : 4 * cmax]).reshape(M * G, cmax, 4)
    std_c[:, :, 1:] *= sc[:, :, None]
    std_c = std_c.reshape(M, G, cmax, 4)
    acc = s.t["acc_total"].sum().item()
    mb = ModelBatch(models=models, M=M, G=G, cmax=cmax, voigt=bool(voigt), sd_free=bool(sd_free),
                    bic=out["bic"], chi2=out["chi2"], lnL=out["lnL"], logp=out["logp"], nm_iters=out["iters"],
                    theta=comps, sd=sd, mean=mean_c, std=std_c, comom=com, ntally=ntal,
                    sd_mean=mean_s[:, 4 * cmax].reshape(M, G) if sd_free else None,
                    sd_std=np.sqrt(var_s[:, 4 * cmax]).reshape(M, G) if sd_free else None,
                    primary_rows=prow.reshape(M, G), accepted_post_burn=int(acc),
                    voigt_evals=s.voigt_evals(iterations), work=s.work(), n_chains=M * per,
                    scale=regions.scale[[r for r, _ in models]], iterations=int(iterations),
                    host_regions=getattr(regions, "host_regions", None))
    if timing is not None:
        timing["host_post_s"] = timing.get("host_post_s", 0.0) + (_time.perf_counter() - _t3)
    if keep_sampler:
        mb.sampler = s
        mb.inv = inv
    return mb


def ladder_examined(n0, tables, chi_limit=1.5):
    """The largest candidate count ``ladder_replay`` LOOKS AT on these tables, i.e. what the reference's sequential
    ladder would have fitted: candidates fitted ahead of the rule (``fit_local``'s speculative window) lie beyond it
    and are left out of a region's record, so records do not depend on how the candidates were scheduled."""
    n = int(n0)
    if n not in tables:
        return n
    bic_old = tables[n][0][-1]
    n += 1
    while n in tables:
        bic_array, red = tables[n]
        cur = float(np.average(bic_array))
        if not bic_old > cur or np.average(red) < chi_limit:
            return n
        bic_old = cur
        n += 1
    return n - 1


def ladder_replay(n0, tables, chi_limit=1.5, verbose=False, log=print):
    """Replay of ``VPregion.region_fit``'s acceptance rule (vamp_1.0/vpregion.py:54-91) on
    precomputed tables ``tables[n] = (bic_array, red_chi_array)``.

    Returns ``(kept_n, exit_n, decided)``; ``decided`` is False when the rule needs a
    candidate that is not in ``tables`` yet (the caller fits more and replays).  Reference
    quirks kept: the first candidate is compared through ``bic_array[-1]`` (not the mean);
    when BIC stops decreasing the previous candidate's fit is kept."""
    n = int(n0)
    if n not in tables:
        return None, n, False
    bic_old = tables[n][0][-1]
    kept = n
    n += 1
    while True:
        if n not in tables:
            return kept, n, False
        bic_array, red = tables[n]
        cur = float(np.average(bic_array))
        if bic_old > cur:
            if verbose:
                log("Old BIC value of {:.2f} is greater than the current {:.2f}.".format(bic_old, cur))
            bic_old, kept = cur, n
            if np.average(red) < chi_limit:
                if verbose:
                    log("Reduced Chi squared is less than {}".format(chi_limit))
                    log("Final n={}".format(n))
                return kept, n, True
            n += 1
            if verbose:
                log("Increasing the number of lines to: {}".format(n))
        else:
            if verbose:
                log("BIC increased with increasing the line number, stopping.")
                log("Final n={}.".format(n - 1))
            return kept, n - 1, True
