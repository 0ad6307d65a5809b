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
C4_PER_GPU = int(os.environ.get("VAMP_BENCH_C4", "512"))      # regions per GPU of the "fitted regions/s" sample (0 = skip)
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


def reference_sample(cores, iters, seed=0):
    """Bounded sample of the C2 workload on ``cores`` host cores: chains for n = 1..6 round-robin."""
    import multiprocessing as mp
    jobs = [(2, 1 + (i % N_CAND), iters, seed + i) for i in range(max(cores, 1))]
    t0 = time.perf_counter()
    if cores > 1:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_ref_worker, jobs)
    else:
        res = [_ref_worker(jobs[0])]
    wall = time.perf_counter() - t0
    evals = sum(r[0] for r in res)
    return evals / wall, evals, wall


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
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def k1_point(ctx, engine, flops, synth, peak_tf, P=256, C=8, B=262144):
    """One likelihood-only launch (K1, fp32): 2^18 parameter vectors x 8 Voigt components x 256 pixels,
    posterior-like parameters.  Algorithmic flops from vamp_b200/flops.py with the path mix of the kernel's own
    pair policy (all 64 pixels of a row pair at |z|^2 >= 32 -> continued fraction) evaluated on the host."""
    import torch
    from vamp_b200 import _lib
    dev = ctx.torch_device
    reg = synth.make_region(P, 8, 50.0, 7000 + P + C, voigt=True)
    regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])])
    rng = np.random.default_rng(5)
    nb = 4096
    th = np.stack([synth.prior_draw_theta(reg["nu"], C, True, rng, posterior_like=reg["truth"]) for _ in range(nb)])
    ths = regs.to_scaled(th, np.zeros(nb, np.int32))
    u = (reg["nu"] - regs.nu0[0]) / regs.scale[0]
    du = (u - u[(P - 1) // 2]).reshape(P // 64, 64)                       # row pairs
    cb = ths[:, :, 1] - u[(P - 1) // 2]
    g, y = 1.66510922231539551270 / ths[:, :, 3], 0.83255461115769775635 * ths[:, :, 2] / ths[:, :, 3]
    smin = (np.abs(du[None, None] - cb[:, :, None, None]).min(axis=3) * g[:, :, None]) ** 2 + (y ** 2)[:, :, None]
    far_frac = float(np.mean(smin >= 32.0))
    d_th = torch.from_numpy(np.tile(engine.pack_theta(ths, None, C), (B // nb, 1))).to(dev)
    d_ir = torch.zeros(B, dtype=torch.int32, device=dev)
    d_in = torch.full((B,), C, dtype=torch.int32, device=dev)
    d_sd = torch.full((B,), 0.02, dtype=torch.float64, device=dev)
    d_lnL = torch.empty(B, dtype=torch.float64, device=dev)
    d_chi = torch.empty(B, dtype=torch.float64, device=dev)

    def run():
        _lib.check(ctx.lib.vamp_loglike_dev(ctx.handle, regs.handle, _lib.VAMP_F32, _lib.VAMP_VOIGT, B, d_ir.data_ptr(),
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
    f = flops.F32
    fl = evals * (far_frac * f["far"] + (1 - far_frac) * f["core"] + f["shape_common"] + 2) + float(P) * B * (f["like"] + 1)
    regs.close()
    return {"workload": "K1 fp32: %d parameter vectors x %d Voigt components x %d px, posterior-like" % (B, C, P),
            "evals_per_sec": evals / t, "kernel_ms": t * 1e3, "far_path_fraction": far_frac,
            "achieved_tflops": fl / t / 1e12, "frac_of_fp32_peak": fl / t / 1e12 / peak_tf,
            "hbm_gbs": B * (P * 12 / 8 + 32 * C + 16) / t / 1e9}


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from vamp_b200 import batch, engine, flops, synth
    from vamp_b200.fit_regions import fit_regions

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
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
    work_per_step = {}

    def device_step(seed):
        mb = batch.fit_models(regs, models, voigt=True, sd_free=True, iterations=ITERS, burn=BURN, thin=THIN,
                              restarts=RESTARTS, chains_per_restart=CHAINS_PER_RESTART, polish_top=4, dtype="f32",
                              seed=seed, region_uids=[region_id], keep_sampler=True)
        return mb

    clocks = ClockSampler(local_rank)
    clocks.start()
    # warm-up (also measures the algorithmic work and the K2 / K3 kernel times with events on the stream)
    k2_ms, k3_ms = [], []
    mb = None
    for w in range(max(args.warmup, 1)):
        flush.fill_(float(w))
        mb = device_step(seed=1000 + w)
        torch.cuda.synchronize()
    work = mb.sampler.work()
    chains = mb.n_chains
    n_free_evals = work["shape_pix"]
    # kernel-level timing of the dominant kernel (K2) on the launching stream
    s = mb.sampler
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    s.init()
    s.t["work"].zero_()
    torch.cuda.synchronize()
    ev[0].record()
    s.run(ITERS)
    ev[1].record()
    torch.cuda.synchronize()
    k2_time = ev[0].elapsed_time(ev[1]) * 1e-3
    work_k2 = s.work()
    flops_k2 = flops.k2_flops(work_k2, "f32", voigt=True)

    del mb, s                              # nothing from the warm-up stays allocated during the timed region
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
    stats = {}
    t0 = time.perf_counter()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    e2e_evals = 0.0
    d2h = 0
    for k in range(args.steps):
        flush.fill_(float(k))
        st = {}
        res = fit_regions(host_regions, uids=[region_id], distributed=False, voigt=True, iterations=ITERS, burn=BURN,
                          thin=THIN, restarts=RESTARTS, chains_per_restart=CHAINS_PER_RESTART, window=N_CAND, n0=[1],
                          nmax=N_CAND, seed=3000 + k, stats=st)
        e2e_evals += st["voigt_evals_actual"] if "voigt_evals_actual" in st else st["voigt_evals"]
        stats = st
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
    # -------- the metric's second half, "fitted regions/s": a bounded sample of BASELINE config C4 (mixed regions,
    #          32-256 px, 1-8 true components, S/N 10/30/100; 3 restarts x 8 chains, ladder window 3) through the public
    #          sharded API - C4_PER_GPU regions per GPU (weak scaling), LPT partition over the ranks, one all-gather of
    #          the result records; host wall clock between barriers, max over ranks (the ladder logic runs on the host)
    c4 = None
    if C4_PER_GPU > 0:
        Rc4 = C4_PER_GPU * world
        c4_regs = [synth.config_c4_region(10_000 + i) for i in range(Rc4)]
        c4_host = [(r["nu"], r["flux"], r["noise"]) for r in c4_regs]
        c4_truth = np.array([len(r["truth"]) for r in c4_regs])
        c4_kw = dict(voigt=True, iterations=ITERS, burn=BURN, thin=THIN, restarts=3, chains_per_restart=8, window=3, nmax=8)
        fit_regions(c4_host[: 32 * world], uids=list(range(10_000, 10_000 + 32 * world)), seed=0, **c4_kw)   # warm-up
        c4_times = []
        for rep in range(3):               # three fits of the sample (own Philox seeds), the median time is reported
            barrier()
            tc = time.perf_counter()
            res4 = fit_regions(c4_host, uids=list(range(10_000, 10_000 + Rc4)), seed=1 + rep, **c4_kw)
            barrier()
            c4_t = torch.tensor([time.perf_counter() - tc], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(c4_t, op=dist.ReduceOp.MAX)
            c4_times.append(float(c4_t[0]))
        c4_s = float(np.median(c4_times))
        nfit = np.array([r.n for r in res4])
        c4 = {"workload": "C4 sample: %d mixed synthetic Voigt regions per GPU (32-256 px, 1-8 true components, S/N 10/30/100), "
                          "ladder window 3, 3 restarts x 8 chains x %d iterations, NM polish, BIC selection; fit_regions from host "
                          "buffers, LPT-sharded over %d rank(s), one all-gather of records" % (C4_PER_GPU, ITERS, world),
              "regions": Rc4, "seconds": c4_s, "regions_per_sec": Rc4 / c4_s, "seconds_all": [round(t, 3) for t in c4_times],
              "n_equals_truth": float(np.mean(nfit == c4_truth)), "n_within_1": float(np.mean(np.abs(nfit - c4_truth) <= 1)),
              "ladder_decided": float(np.mean([r.status == 0 for r in res4]))}
    clk = clocks.stop(t_host0, time.perf_counter())       # samples of the timed regions (device step, e2e, C4 sample)
    h2d = int(sum(a.nbytes for a in host_regions[0]) + 4)
    rec_bytes = int(8 * (6 + 2 * N_CAND * RESTARTS + 1 + 3 * (1 + N_CAND) * 4 + 3))

    if rank != 0:
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
    achieved_tf = flops_k2 / k2_time / 1e12
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # algorithmic HBM bytes of one K2 launch: spectrum tile per chain + chain state in/out
    P = 256
    ts = 4 * N_CAND + 1
    # spectrum tile (du, flux: 8 B/px) once per CTA of 4 chains + chain state in and out
    alg_bytes = chains * (P * 8 // 4 + 2 * (ts * (8 * 4 + 4 * 2 + 4 * 3) + 6 * N_CAND * 8 + 64))
    # dram__bytes_read.sum + dram__bytes_write.sum of this launch (same grid, 3000 iterations) from one
    # `ncu --set full` capture of `bench.py --steps 1 --warmup 1`: profiles/r1_k2_bench_launch_full.txt
    K2_DRAM_TRAFFIC_BYTES = 34.698496e6 + 3.837952e6
    roofline = {"bound": "fp32", "kernel": "k2_mcmc_kernel<float, VOIGT>", "achieved": achieved_tf, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "peak_source": "SMs x 128 x 2 x sm_max_mhz (MEASURED_PEAKS.json %s MHz, %d SMs)" % (sm_max, nsm),
                "kernel_ms": k2_time * 1e3, "algorithmic_flops_per_launch": flops_k2,
                "evals_per_launch": work_k2["shape_pix"], "far_path_fraction": work_k2["far_pix"] / max(work_k2["shape_pix"], 1),
                "kernel_evals_per_sec": work_k2["shape_pix"] / k2_time,
                "traffic": K2_DRAM_TRAFFIC_BYTES, "traffic_unit": "bytes per launch (ncu dram read + write)",
                "hbm": {"achieved_gbs": alg_bytes / k2_time / 1e9, "peak_gbs": hbm_peak,
                        "algorithmic_bytes_per_launch": alg_bytes}}
    if clk.get("sm_mhz"):
        roofline["frac_at_observed_clock"] = achieved_tf / flops.fp32_peak_tflops(nsm, clk["sm_mhz"])
    # measured FMA-chain burst on this pool's B200s (tools/micro/ffma2_bench.cu: FFMA 72.3, FFMA2 74.0 TFLOP/s)
    roofline["frac_of_measured_ffma2_burst"] = achieved_tf / 74.0

    # -------- K1, the fused likelihood kernel on its own (BASELINE config C5, one point): the same Voigt arithmetic
    #          without the Metropolis bookkeeping - the figure to hold against the FP32 roofline
    k1 = k1_point(ctx, engine, flops, synth, peak_tf)

    cpu_v, cpu_e, cpu_w = reference_sample(1, 60, seed=7)
    cpu_baseline = {"value": cpu_v, "unit": UNIT, "cores": 1, "kind": "port",
                    "sample": "restated PyMC2 scalar Metropolis over the reference density (numpy + scipy wofz), C2 region, "
                              "n=1, 60 iterations on one core (%.1f s)" % cpu_w}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(world),
            "clocks": clk,
            "e2e": {"value": e2e_evals / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": rec_bytes,
                    "ms_per_step": 1e3 * e2e_s / max(args.steps, 1)},
            "gpu_launches": 4 * args.steps,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "k1_likelihood": k1,
            "fitted_regions": c4,
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
