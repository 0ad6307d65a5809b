#!/usr/bin/env python
"""bench.py - the driver's measurement contract for the VAMP hot path on B200.

Metric (BASELINE.json): Voigt logL pixel-component evaluations per second, measured over whole
fits of the configuration the metric is quoted on (configs[1], "C2"):

    one synthetic blended 256-pixel region per GPU, BIC model-selection sweep over n = 1..6
    Voigt components, 4096 chains per candidate (4 restart groups x 1024), 3000 iterations
    (burn 300, thin 15) of the scalar Metropolis sampler, Nelder-Mead MAP polish, BIC table,
    component-count selection.

A *step* = one such fit.  ``value`` counts the Voigt pixel-component evaluations the sampler
actually computed (device counters: one component re-evaluated over the region per in-bounds
centroid/width step; amplitude, noise and out-of-bounds steps compute none) divided by the
device time of the step, inputs resident in HBM.  ``e2e`` times the public API call
(``vamp_b200.fit_regions``) with host numpy buffers: spectrum upload, fit, read-back of the
per-region result records.  N > 1: one rank per GPU, each fits its own region (weak scaling,
Philox keyed by global region id), one NCCL all-gather of the result records per step.

Further keys of the JSON line: ``precisions`` (the same K2 launch and a K1 launch in float64 - BASELINE config C5 is
"fp32 vs fp64"), ``fitted_regions`` (the metric's second half: a FIXED total of BASELINE C4 regions sharded over the N
GPUs - strong scaling - with a world-size independent checksum of the result records), ``fitted_regions_weak``,
``c3_sightline`` (BASELINE config C3), ``c1_region_fit`` (BASELINE config C1 on the GPU beside the CPU port) and
``ladder_vs_oracle`` (chosen-n agreement with the restated reference fit on 32 regions).

``--impl reference``: the reference's arithmetic on the host cores - the oracle's restatement
of the PyMC 2 scalar Metropolis sampler over the reference density (numpy + scipy.special.wofz,
exactly what vamp_1.0/vpfits.py computes through astropy/PyMC), one process per core, on a
bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "voigt_logl_pixel_component_evals_per_sec"
UNIT = "evals/s"
N_CAND = 6
RESTARTS = 4
CHAINS_PER_RESTART = 1024
ITERS, BURN, THIN = 3000, 300, 15
C4_TOTAL = int(os.environ.get("VAMP_BENCH_C4_TOTAL", "16384"))  # "fitted regions/s": FIXED total sharded over the N GPUs (strong scaling;
#                                                                  100000 = BASELINE config C4 at full size; 0 = skip)
C4_PER_GPU = int(os.environ.get("VAMP_BENCH_C4", "512"))        # the weak-scaling sample of round 1: regions per GPU (0 = skip)
# ladder window (2, 1): the first round fits the two candidates the first BIC comparison needs, every later round one
# more - no speculative fits, like the reference's sequential ladder (round 1 fitted three candidates per round; the
# ladder's decisions are the same, chains being keyed by (region, n): 479 -> 731 regions/s on 4096 regions)
C4_KW = dict(voigt=True, iterations=3000, burn=300, thin=15, restarts=3, chains_per_restart=8, window=(2, 1), nmax=8)
WORKLOAD = ("C2: 256-px synthetic blended Voigt region (3 true components, S/N 50) per GPU; BIC sweep n=1..6, "
            "4096 chains per candidate (4 restarts x 1024), 3000 iterations burn 300 thin 15, NM polish, BIC selection")


def config_dict(n_gpus):
    return {"workload": WORKLOAD, "pixels": 256, "candidates": "1..6", "chains_per_candidate": RESTARTS * CHAINS_PER_RESTART,
            "iterations": ITERS, "burn": BURN, "thin": THIN, "regions_per_gpu": 1, "parallelism": "regions sharded, %d rank(s); weak scaling: one C2 region per GPU (same spectrum, own chains)" % n_gpus,
            "l2": "flushed between steps (512 MiB write); working set is compute-bound and far below L2"}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Statistics over the samples that arrived in [t0, t1] (host perf_counter): the sampler is started
        before the warm-up so that nvidia-smi's start-up (which can stall kernel submission for 0.1-0.3 s)
        never falls inside a timed region."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        rows = [r for t, r in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1 + 0.3)]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = [v for v in sm if v > 500] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ reference arm
def _ref_worker(args):
    """One host core: restated PyMC 2 scalar Metropolis on the reference density."""
    region_id, n, iters, seed = args
    from oracle import sampler as OS
    from vamp_b200 import synth
    reg = synth.config_c2(region_id=region_id)
    model = OS.Model(reg["nu"], reg["flux"], n, True)
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    OS.metropolis_sample(model, model.initial(rng), iters, iters // 10, THIN, rng)
    dt = time.perf_counter() - t0
    # the restated sampler recomputes every component at every scalar step (no LazyFunction cache)
    evals = iters * model.k * n * len(reg["nu"])
    return evals, dt


def reference_sample(cores, iters, seed=0, n_fixed=None):
    """Bounded sample of the C2 workload on ``cores`` host cores: chains for n = 1..6 round-robin (``n_fixed``: all
    chains of that component count)."""
    import multiprocessing as mp
    jobs = [(2, n_fixed or (1 + (i % N_CAND)), iters, seed + i) for i in range(max(cores, 1))]
    t0 = time.perf_counter()
    if cores > 1:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_ref_worker, jobs)
    else:
        res = [_ref_worker(jobs[0])]
    wall = time.perf_counter() - t0
    evals = sum(r[0] for r in res)
    return evals / wall, evals, wall


def c1_cpu_region_fit():
    """BASELINE config C1 on the host: ``VPregion.region_fit`` (vamp_1.0/vpregion.py:42-91) restated - the oracle's
    PyMC 2 sampler + Nelder-Mead MAP + BIC ladder on the one-component 200-px HI region, one core (BASELINE.md
    section 3.2: wall time, scalar steps/s, final n beside the GPU figure).  The third-party sampler is a port."""
    from oracle import region as OR
    from oracle import sampler as OS
    from vamp_b200 import synth
    reg = synth.config_c1()
    n0 = OR.estimate_n(reg["flux"])
    freedom = OR.freedom(len(reg["flux"]), n0)
    rng = np.random.default_rng(1)
    steps = [0]

    def fb(n):
        b, r, _ = OS.find_bic(reg["nu"], reg["flux"], n, reg["noise"], freedom, voigt=True, iterations=ITERS, rng=rng)
        steps[0] += 3 * ITERS * (4 * n + 1)
        return b, r

    t0 = time.perf_counter()
    kept, exit_n, hist = OR.region_fit_ladder(n0, fb, chi_limit=1.5, n_cap=n0 + 3)
    wall = time.perf_counter() - t0
    return {"wall_s": wall, "scalar_steps": steps[0], "scalar_steps_per_s": steps[0] / wall, "final_n": int(kept),
            "candidates": [int(h[0]) for h in hist], "bic": [round(float(h[1]), 2) for h in hist], "cores": 1,
            "kind": "port", "what": "restated VPregion.region_fit: 3 restarts x (%d iterations PyMC 2 scalar Metropolis, "
            "Nelder-Mead MAP, BIC) per candidate n, numpy + scipy wofz" % ITERS}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    iters = 150
    for _ in range(args.warmup if args.warmup < 2 else 1):
        reference_sample(cores, 5)
    t_steps, ev = [], 0
    for k in range(args.steps):
        v, e, w = reference_sample(cores, iters, seed=100 + k)
        t_steps.append(w)
        ev += e
    total = sum(t_steps)
    value = ev / total
    sample = ("restated PyMC2 scalar Metropolis over the reference density (numpy + scipy.special.wofz), C2 region, "
              "%d chains (n=1..6 round-robin) x %d iterations per step, one process per core" % (cores, iters))
    cfg = config_dict(args.gpus)
    # what THIS arm ran (a bounded sample of the configuration above; the metric is a rate)
    cfg["reference_sample"] = {"chains_per_step": cores, "iterations_per_chain": iters, "candidates": "1..6 round-robin",
                               "processes": cores, "note": "the GPU arm's 4096 chains x 3000 iterations per candidate would "
                               "take hours on the host; every step here is %d chains x %d iterations" % (cores, iters)}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def k1_point(ctx, engine, flops, synth, peak_tf, P=256, C=8, B=262144, dtype="f32", prior=False):
    """One likelihood-only launch (K1): B parameter vectors x 8 Voigt components x 256 pixels, posterior-like
    parameters (``prior``: prior draws instead - SURVEY 8d asks for both, the w(z) branch mix differs).  Algorithmic flops from vamp_b200/flops.py with the path mix of the kernel's own pair policy (all 64
    pixels of a row pair at |z|^2 >= R2 -> continued fraction) evaluated on the host.  ``dtype`` "f64": the float64
    Faddeeva (Weideman N = 32), against the FP64 peak."""
    import torch
    from vamp_b200 import _lib
    dev = ctx.torch_device
    f64 = dtype == "f64"
    reg = synth.make_region(P, 8, 50.0, 7000 + P + C, voigt=True)
    regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])])
    rng = np.random.default_rng(5)
    nb = 4096
    th = np.stack([synth.prior_draw_theta(reg["nu"], C, True, rng, posterior_like=None if prior else reg["truth"])
                   for _ in range(nb)])
    if prior:
        th[:, :, 2:] = np.maximum(th[:, :, 2:], 1e-3 * (reg["nu"][1] - reg["nu"][0]))      # widths > 0 (the prior's open bound)
    ths = regs.to_scaled(th, np.zeros(nb, np.int32))
    u = (reg["nu"] - regs.nu0[0]) / regs.scale[0]
    du = (u - u[(P - 1) // 2]).reshape(P // 64, 64)                       # row pairs
    cb = ths[:, :, 1] - u[(P - 1) // 2]
    g, y = 1.66510922231539551270 / ths[:, :, 3], 0.83255461115769775635 * ths[:, :, 2] / ths[:, :, 3]
    smin = (np.abs(du[None, None] - cb[:, :, None, None]).min(axis=3) * g[:, :, None]) ** 2 + (y ** 2)[:, :, None]
    far_frac = float(np.mean(smin >= (64.0 if f64 else 32.0)))
    d_th = torch.from_numpy(np.tile(engine.pack_theta(ths, None, C), (B // nb, 1))).to(dev)
    d_ir = torch.zeros(B, dtype=torch.int32, device=dev)
    d_in = torch.full((B,), C, dtype=torch.int32, device=dev)
    d_sd = torch.full((B,), 0.02, dtype=torch.float64, device=dev)
    d_lnL = torch.empty(B, dtype=torch.float64, device=dev)
    d_chi = torch.empty(B, dtype=torch.float64, device=dev)
    code = _lib.VAMP_F64 if f64 else _lib.VAMP_F32

    def run():
        _lib.check(ctx.lib.vamp_loglike_dev(ctx.handle, regs.handle, code, _lib.VAMP_VOIGT, B, d_ir.data_ptr(),
                                            d_in.data_ptr(), C, d_th.data_ptr(), 4 * C + 1, d_sd.data_ptr(), d_lnL.data_ptr(),
                                            d_chi.data_ptr(), None, P))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = float(np.median(ms)) * 1e-3
    evals = float(P) * C * B
    f = flops.F64 if f64 else flops.F32
    fl = evals * (far_frac * f["far"] + (1 - far_frac) * f["core"] + f["shape_common"] + 2) + float(P) * B * (f["like"] + 1)
    regs.close()
    return {"workload": "K1 %s: %d parameter vectors x %d Voigt components x %d px, %s" % (
                dtype, B, C, P, "prior draws" if prior else "posterior-like"),
            "evals_per_sec": evals / t, "kernel_ms": t * 1e3, "far_path_fraction": far_frac,
            "achieved_tflops": fl / t / 1e12, ("frac_of_fp64_peak" if f64 else "frac_of_fp32_peak"): fl / t / 1e12 / peak_tf,
            "hbm_gbs": B * (P * (24 if f64 else 12) / 8 + 32 * C + 16) / t / 1e9}


def cpu_likelihood_sample(synth, P=256, C=8, seconds=2.5):
    """SURVEY 8d, CPU baseline (1): likelihood only - the oracle's reference-line path (numpy + scipy.special.wofz, what
    vpfits.py:58-76, 334-341 compute through astropy / PyMC) on ONE host core, K1's workload shape, for ~``seconds``."""
    from oracle import density as D
    reg = synth.make_region(P, 8, 50.0, 7000 + P + C, voigt=True)
    rng = np.random.default_rng(5)
    ths = [synth.prior_draw_theta(reg["nu"], C, True, rng, posterior_like=reg["truth"]) for _ in range(64)]
    n, t0, acc = 0, time.perf_counter(), 0.0
    while time.perf_counter() - t0 < seconds:
        th = ths[n % len(ths)]
        acc += D.normal_loglike(reg["flux"], D.model_flux(reg["nu"], th, True), sd=0.02)
        n += 1
    dt = time.perf_counter() - t0
    return {"evals_per_sec": n * P * C / dt, "cores": 1, "kind": "port", "vectors": n, "seconds": dt,
            "what": "oracle density: exp(-sum of wofz Voigt profiles) + Normal lnL, %d components x %d px per vector" % (C, P)}


# ------------------------------------------------------------------------------------------ GPU arm
def _c4_hosts(first_uid, count, procs):
    """``count`` BASELINE C4 regions ``(nu, flux, noise)`` + their true component counts, generated on forked host
    processes (called before this process touches CUDA)."""
    from vamp_b200 import synth
    uids = list(range(first_uid, first_uid + count))
    if count >= 2048 and procs > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            regs = pool.map(synth.config_c4_region, uids, chunksize=256)
    else:
        regs = [synth.config_c4_region(u) for u in uids]
    return [(r["nu"], r["flux"], r["noise"]) for r in regs], np.array([len(r["truth"]) for r in regs]), uids


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from vamp_b200 import batch, engine, flops, synth
    from vamp_b200.fit_regions import fit_regions

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # workload generation first: host processes are forked before this process touches CUDA
    procs = max(1, min(16, (os.cpu_count() or 8) // max(world, 1)))
    c4s_host = c4w_host = None
    if C4_TOTAL > 0:
        c4s_host, c4s_truth, c4s_uids = _c4_hosts(10_000, max(C4_TOTAL, C4_PER_GPU * world), procs)
    elif C4_PER_GPU > 0:
        c4s_host, c4s_truth, c4s_uids = _c4_hosts(10_000, C4_PER_GPU * world, procs)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the GPU arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = engine.default_context(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # weak scaling, fixed per-GPU work: every GPU fits the SAME C2 spectrum (region 2) with its own chains - the
    # Philox streams are keyed by the global uid 2 + rank - so the max over ranks measures the machine, not the
    # difference between one synthetic region and another (region 2 vs 9 differ by ~5 % in accepted work)
    region_id = 2 + rank
    reg = synth.config_c2(region_id=2)
    host_regions = [(reg["nu"], reg["flux"], reg["noise"])]
    models = [(0, n) for n in range(1, N_CAND + 1)]
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    regs = engine.Regions(ctx, host_regions)

    # -------- device-resident step: prior draw -> K2 -> top-K -> K3 -> BIC table (no host copies)
    def device_step(seed):
        mb = batch.fit_models(regs, models, voigt=True, sd_free=True, iterations=ITERS, burn=BURN, thin=THIN,
                              restarts=RESTARTS, chains_per_restart=CHAINS_PER_RESTART, polish_top=4, dtype="f32",
                              seed=seed, region_uids=[region_id], keep_sampler=True)
        return mb

    clocks = ClockSampler(local_rank)
    clocks.start()
    # warm-up (also measures the algorithmic work and the K2 kernel time with events on the stream)
    mb = None
    for w in range(max(args.warmup, 1)):
        flush.fill_(float(w))
        mb = device_step(seed=1000 + w)
        torch.cuda.synchronize()
    chains = mb.n_chains
    # kernel-level timing of the dominant kernel (K2) on the launching stream: fp32 (the headline) and float64
    s = mb.sampler

    def time_k2(smp):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        smp.init()
        smp.t["work"].zero_()
        torch.cuda.synchronize()
        ev[0].record()
        smp.run(ITERS)
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) * 1e-3, smp.work()

    k2_time, work_k2 = time_k2(s)
    flops_k2 = flops.k2_flops(work_k2, "f32", voigt=True)
    s64 = engine.Sampler(regs, s.item_region_h, s.item_ncomp_h, s.t["chain_uid"].cpu().numpy().view(np.uint64), voigt=True,
                         sd_free=True, dtype="f64", seed=77, burn=BURN, thin=THIN, tune_interval=50, tune_throughout=True,
                         cmax=s.cmax)
    k2_time64, work_k2_64 = time_k2(s64)
    flops_k2_64 = flops.k2_flops(work_k2_64, "f64", voigt=True)
    del mb, s, s64                         # nothing from the warm-up stays allocated during the timed region
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e0.record()
    evals = 0.0
    last = None
    step_ev = [e0]
    for k in range(args.steps):
        flush.fill_(float(k))
        last = None                        # release the previous fit first: the caching allocator then reuses its
        last = device_step(seed=2000 + k)  # blocks instead of growing (a cudaMalloc inside the timed region)
        evals += last.work["shape_pix"]
        step_ev.append(torch.cuda.Event(enable_timing=True))
        step_ev[-1].record()
        if world > 1:                      # gather the BIC tables (the only data-path exchange)
            t = torch.from_numpy(np.ascontiguousarray(last.bic)).to(dev)
            out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=dev)
            dist.all_gather_into_tensor(out, t)
    e1.record()
    barrier()
    dev_s = e0.elapsed_time(e1) * 1e-3
    t_host1 = time.perf_counter()
    wall_s = t_host1 - t_host0
    tt = torch.tensor([dev_s, evals], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_s, evals = float(tmax[0]), float(tsum[1])
    value = evals / dev_s

    # -------- end to end through the public API with host buffers
    barrier()
    t0 = time.perf_counter()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    e2e_evals = 0.0
    for k in range(args.steps):
        flush.fill_(float(k))
        st = {}
        res = fit_regions(host_regions, uids=[region_id], distributed=False, voigt=True, iterations=ITERS, burn=BURN,
                          thin=THIN, restarts=RESTARTS, chains_per_restart=CHAINS_PER_RESTART, window=N_CAND, n0=[1],
                          nmax=N_CAND, seed=3000 + k, stats=st)
        e2e_evals += st["voigt_evals_actual"] if "voigt_evals_actual" in st else st["voigt_evals"]
        rec_bytes = int(res.records.nbytes)
        if world > 1:
            t = torch.tensor([float(res[0].n)], dtype=torch.float64, device=dev)
            out = torch.empty(world, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(out, t)
    ee1.record()
    barrier()
    e2e_s = max(ee0.elapsed_time(ee1) * 1e-3, time.perf_counter() - t0 if world == 1 else 0.0)
    e2e_t = torch.tensor([e2e_s, e2e_evals], dtype=torch.float64, device=dev)
    if world > 1:
        a = e2e_t.clone()
        dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = e2e_t.clone()
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e2e_s, e2e_evals = float(a[0]), float(b[1])

    # -------- the metric's second half, "fitted regions/s" (BASELINE config C4: mixed regions, 32-256 px, 1-8 true
    #          components, S/N 10/30/100; 3 restarts x 8 chains, ladder window 3) through the public sharded API:
    #          host buffers in, LPT partition over the ranks, one all-gather of the result records; host wall clock
    #          between barriers, max over ranks (the ladder logic runs on the host).
    #          fitted_regions      : a FIXED total sharded over the N GPUs (strong scaling), one timed fit
    #          fitted_regions_weak : C4_PER_GPU regions per GPU (weak scaling, round 1's sample), median of 3 fits
    def c4_fit(host, truth, uids, reps, seed0):
        times, res4, st4 = [], None, {}
        for rep in range(reps):
            st4 = {}
            barrier()
            tc = time.perf_counter()
            res4 = fit_regions(host, uids=uids, seed=seed0 + rep, stats=st4, **C4_KW)
            barrier()
            times.append(max_over_ranks(time.perf_counter() - tc))
        nfit = res4.n
        tm = st4.get("timing", {})
        return {"regions": len(host), "seconds": float(np.median(times)), "regions_per_sec": len(host) / float(np.median(times)),
                "seconds_all": [round(t, 3) for t in times],
                "n_equals_truth": float(np.mean(nfit == truth)), "n_within_1": float(np.mean(np.abs(nfit - truth) <= 1)),
                "ladder_decided": float(np.mean(res4.status == 0)), "records_sha256": res4.checksum(),
                "rank0_device_ms": {"k2": round(tm.get("k2_ms", 0.0), 1), "polish": round(tm.get("polish_ms", 0.0), 1)},
                # the sampler's rate on this workload (rank 0's shard): shape pixels it evaluated / its K2 device time
                "rank0_k2_evals_per_sec": (st4.get("voigt_evals_actual", 0.0) / (tm["k2_ms"] * 1e-3)) if tm.get("k2_ms") else None,
                "rank0_chains": int(st4.get("chains", 0)),
                "rank0_host_s": {k: round(tm.get(k, 0.0), 3) for k in ("upload_s", "host_setup_s", "host_launch_s",
                                                                       "host_wait_s", "host_post_s", "ladder_s", "records_s")}}

    c4 = c4w = None
    c4_desc = ("mixed synthetic Voigt regions (32-256 px, 1-8 true components, S/N 10/30/100), ladder window (2, 1), 3 restarts x "
               "8 chains x %d iterations, NM polish, BIC selection; fit_regions from host buffers, LPT-sharded over %d "
               "rank(s), one all-gather of records" % (ITERS, world))
    if c4s_host is not None:
        fit_regions(c4s_host[: 64 * world], uids=c4s_uids[: 64 * world], seed=0, **C4_KW)            # warm-up
    if C4_TOTAL > 0:
        c4 = c4_fit(c4s_host[:C4_TOTAL], c4s_truth[:C4_TOTAL], c4s_uids[:C4_TOTAL], 1, 1)
        c4["workload"] = "C4, fixed total: %d %s" % (C4_TOTAL, c4_desc)
        c4["scaling"] = "strong"
    if C4_PER_GPU > 0:
        Rw = C4_PER_GPU * world
        c4w = c4_fit(c4s_host[:Rw], c4s_truth[:Rw], c4s_uids[:Rw], 3, 1)
        c4w["workload"] = "C4 sample: %d per GPU, %s" % (C4_PER_GPU, c4_desc)
        c4w["scaling"] = "weak"
    clk = clocks.stop(t_host0, time.perf_counter())       # samples of the timed regions (device step, e2e, C4 fits)
    h2d = int(sum(a.nbytes for a in host_regions[0]) + 4)

    if rank != 0:
        barrier()                          # rank 0 finishes its single-GPU extras (C3, C1, K1 points, CPU baselines)
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    nsm = torch.cuda.get_device_properties(dev).multi_processor_count
    peak_tf = flops.fp32_peak_tflops(nsm, sm_max)
    peak_tf64 = flops.fp64_peak_tflops(nsm, sm_max)
    achieved_tf = flops_k2 / k2_time / 1e12
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # algorithmic HBM bytes of one K2 launch: spectrum tile per chain + chain state in/out
    P = 256
    ts = 4 * N_CAND + 1
    # spectrum tile (du, flux: 8 B/px) once per CTA of 4 chains + chain state in and out
    alg_bytes = chains * (P * 8 // 4 + 2 * (ts * (8 * 4 + 4 * 2 + 4 * 3) + 6 * N_CAND * 8 + 64))
    # dram__bytes_read.sum + dram__bytes_write.sum of this launch (same grid, 3000 iterations) and the executed FP32
    # instruction counters: profiles/r2_k2.json, written by tools/capture_k2.py from an ncu counter capture
    traffic, traffic_src, executed = None, "profiles/r2_k2.json missing (run tools/capture_k2.py)", None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "r2_k2.json")))
        traffic, traffic_src = cap["traffic_bytes"], "profiles/r2_k2.json (ncu dram__bytes_read.sum + dram__bytes_write.sum)"
        executed = {"source": "profiles/r2_k2.json (ncu counters of the same 3000-iteration launch, tools/capture_k2.py)",
                    "scalar_thread_inst_ffma_fmul_fadd": [cap.get("thread_inst_ffma"), cap.get("thread_inst_fmul"), cap.get("thread_inst_fadd")],
                    "warp_instructions": (cap.get("metrics") or {}).get("smsp__inst_executed.sum"),
                    "fp32_flops_per_warp_lane_instruction": (cap.get("mix") or {}).get("fp32_flops_per_warp_lane_instruction"),
                    "executed_fp32_flops": cap.get("executed_fp32_flops"),
                    "algorithmic_flops_same_launch": cap.get("algorithmic_flops"),
                    "executed_over_algorithmic": cap.get("executed_over_algorithmic")}
    except Exception:
        pass
    roofline = {"bound": "fp32", "kernel": "k2_mcmc_kernel<float, VOIGT>", "achieved": achieved_tf, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "peak_source": "SMs x 128 x 2 x sm_max_mhz (MEASURED_PEAKS.json %s MHz, %d SMs)" % (sm_max, nsm),
                "kernel_ms": k2_time * 1e3, "algorithmic_flops_per_launch": flops_k2,
                "evals_per_launch": work_k2["shape_pix"], "far_path_fraction": work_k2["far_pix"] / max(work_k2["shape_pix"], 1),
                "kernel_evals_per_sec": work_k2["shape_pix"] / k2_time,
                "traffic": traffic, "traffic_unit": "bytes per launch", "traffic_source": traffic_src,
                "executed_vs_algorithmic": executed,
                "hbm": {"achieved_gbs": alg_bytes / k2_time / 1e9, "peak_gbs": hbm_peak,
                        "algorithmic_bytes_per_launch": alg_bytes}}
    if clk.get("sm_mhz"):
        roofline["frac_at_observed_clock"] = achieved_tf / flops.fp32_peak_tflops(nsm, clk["sm_mhz"])
    # measured FMA-chain burst on this pool's B200s (tools/micro/ffma2_bench.cu: FFMA 72.3, FFMA2 74.0 TFLOP/s)
    roofline["frac_of_measured_ffma2_burst"] = achieved_tf / 74.0

    # -------- BASELINE config C5 is "fp32 vs fp64": the same K2 launch (same chain set, 3000 iterations) and a K1
    #          launch in float64 (Weideman N = 32 Faddeeva), against the FP64 peak SMs x 64 x 2 x clock
    k1 = k1_point(ctx, engine, flops, synth, peak_tf)
    k1p = k1_point(ctx, engine, flops, synth, peak_tf, prior=True)
    k1["prior_draws"] = {k: k1p[k] for k in ("evals_per_sec", "kernel_ms", "far_path_fraction", "achieved_tflops", "frac_of_fp32_peak")}
    if world == 1:
        k1["cpu_port"] = cpu_likelihood_sample(synth)
    k1_64 = k1_point(ctx, engine, flops, synth, peak_tf64, B=65536, dtype="f64")
    precisions = {
        "k2_f64": {"workload": "the roofline's K2 launch in float64: %d chains x %d iterations" % (chains, ITERS),
                   "kernel_ms": k2_time64 * 1e3, "evals_per_sec": work_k2_64["shape_pix"] / k2_time64,
                   "achieved_tflops": flops_k2_64 / k2_time64 / 1e12, "frac_of_fp64_peak": flops_k2_64 / k2_time64 / 1e12 / peak_tf64,
                   "fp64_peak_tflops": peak_tf64, "far_path_fraction": work_k2_64["far_pix"] / max(work_k2_64["shape_pix"], 1)},
        "k1_f64": k1_64,
        "k2_f32_over_f64": k2_time64 / k2_time}

    # -------- BASELINE config C3: a SIMBA-like synthetic sightline in HI 1215, CIV 1548/1550 and MgII 2796/2803
    #          (doublets with tied v / b / N), 1000 px per ion, Gaussian LSF (FWHM 2 px, 7 taps), through
    #          VPspectrum.fit_spectrum (detection, pairing, tied fit, harvest) on this GPU
    from vamp_b200 import vpspectrum as VS

    def c3_once(seed):
        t0_ = time.perf_counter()
        nreg, nlines, chi = 0, 0, []
        for sight in synth.config_c3():
            sp = VS.VPspectrum(sight["line"], voigt=True, doublet=sight["doublet"], lsf=sight["lsf"], chi_limit=1.5,
                               convergence_attempts=2, iterations=2000, burn=300, thin=5, restarts=3, chains_per_restart=32,
                               window=2, nmax=3, seed=seed)
            sp.set_spectrum(sight["wave"], sight["flux"], sight["noise"])
            sp.fit_spectrum()
            nreg += len(sp.region_pixels)
            nlines += len(sp.params["b"])
            chi += [float(c) for c in sp.flux_model["chi_squared"]]
        torch.cuda.synchronize()
        return time.perf_counter() - t0_, nreg, nlines, chi

    c3_once(1)
    c3_s, c3_reg, c3_lines, c3_chi = c3_once(3)
    c3 = {"workload": "C3: one synthetic sightline in HI 1215 + CIV 1548/1550 + MgII 2796/2803 (tied doublets), 3 x 1000 px, "
                      "Gaussian LSF FWHM 2 px / 7 taps, S/N 70, VPspectrum.fit_spectrum on 1 GPU (fp32 sampler, 3 restarts x 32 "
                      "chains x 2000 iterations, float64 polish)",
          "seconds": c3_s, "regions": c3_reg, "absorbers_found": c3_lines, "absorbers_true": 9,
          "regions_per_sec": c3_reg / c3_s, "max_reduced_chi2": max(c3_chi) if c3_chi else None}

    # -------- BASELINE config C1: one HI region, 1 Voigt component, 200 px - VPregion.region_fit on the GPU beside the
    #          restated reference fit on one host core
    from vamp_b200.vpregion import VPregion
    c1 = synth.config_c1()

    def c1_gpu():
        t0_ = time.perf_counter()
        rg = VPregion(c1["nu"], c1["flux"], c1["noise"], voigt=True, n_chains=96, seed=9)
        rg.region_fit(verbose=False)
        torch.cuda.synchronize()
        return time.perf_counter() - t0_, rg

    c1_gpu()
    c1_s, rg = c1_gpu()
    # the CPU legs run on rank 0 at N = 1 only (the other ranks of a multi-GPU run would idle for a minute)
    c1_cpu = c1_cpu_region_fit() if world == 1 else {"wall_s": None, "skipped": "CPU baselines run at N = 1 only"}
    c1_line = {"workload": "C1: synthetic HI Ly-alpha region, 1 Voigt component, 200 px, S/N 50; region_fit (BIC ladder, 3 restarts x "
                           "%d iterations per candidate)" % ITERS,
               "gpu_wall_s": c1_s, "gpu_final_n": int(len(rg.fit.estimated_profiles)), "gpu_chains_per_restart": 32,
               "cpu_port": c1_cpu, "gpu_over_cpu_wall": (c1_cpu["wall_s"] / c1_s) if c1_cpu["wall_s"] else None,
               "note": "the reference's estimate_n counts every local minimum of the smoothed flux once there are four: on "
                       "this 200-px, mostly continuum region it starts the ladder at n0 = 11 (noise), on the GPU as on the CPU"}

    # -------- chosen-n agreement with the restated reference fit (tests/golden/ladder_oracle.json: 32 C4 regions fitted
    #          by oracle/make_ladder_golden.py in the build container) - the comparator for n_equals_truth above
    ladder = None
    try:
        gold = [g for g in json.load(open(os.path.join(ROOT, "tests", "golden", "ladder_oracle.json")))["regions"]
                if g["name"].startswith("c4/")]
        gu = [int(g["name"].split("/")[1]) for g in gold]
        gh = [synth.config_c4_region(u) for u in gu]
        rs = fit_regions([(r["nu"], r["flux"], r["noise"]) for r in gh], uids=gu, seed=11, distributed=False, **C4_KW)
        ng, no = rs.n, np.array([g["kept_n"] for g in gold])
        tr = np.array([g["truth_n"] for g in gold])
        ladder = {"regions": len(gold), "gpu_n_equals_truth": float(np.mean(ng == tr)), "oracle_n_equals_truth": float(np.mean(no == tr)),
                  "gpu_n_within_1": float(np.mean(np.abs(ng - tr) <= 1)), "oracle_n_within_1": float(np.mean(np.abs(no - tr) <= 1)),
                  "gpu_equals_oracle": float(np.mean(ng == no)), "gpu_within_1_of_oracle": float(np.mean(np.abs(ng - no) <= 1)),
                  "oracle": "restated VPregion.region_fit / find_bic (1 chain, PyMC 2 tune interval 1000, MAP from the last state)"}
    except Exception as exc:                                   # the fixture is optional for the bench
        ladder = {"unavailable": str(exc)}

    if world == 1:
        cpu_v, cpu_e, cpu_w = reference_sample(1, 6000, seed=7, n_fixed=3)
        cpu_baseline = {"value": cpu_v, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": "restated PyMC2 scalar Metropolis over the reference density (numpy + scipy wofz), C2 region, "
                                  "one chain of the n=3 candidate, 6000 iterations on one core (%.1f s); the restated sampler "
                                  "recomputes all n components per scalar step and is credited with them" % cpu_w}
    else:
        cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "timed at N = 1 only"}
    barrier()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(world),
            "clocks": clk,
            "e2e": {"value": e2e_evals / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": rec_bytes,
                    "ms_per_step": 1e3 * e2e_s / max(args.steps, 1)},
            "gpu_launches": 4 * args.steps,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "k1_likelihood": k1, "precisions": precisions,
            "fitted_regions": c4, "fitted_regions_weak": c4w, "c3_sightline": c3, "c1_region_fit": c1_line,
            "ladder_vs_oracle": ladder,
            "c2_regions_per_sec": world * args.steps / dev_s,
            "chains_per_step": chains * world,
            "selected_n": int(batch.ladder_replay(1, {n: (last.bic[n - 1], last.chi2[n - 1] / 253.0) for n in range(1, N_CAND + 1)})[0]),
            "host_wall_s": wall_s,
            "step_ms": [round(step_ev[k].elapsed_time(step_ev[k + 1]), 2) for k in range(args.steps)]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="vamp_b200", choices=["vamp_b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
