"""Reproducible roofline evidence for K2 (the kernel ``bench.py``'s roofline is quoted on).

    python tools/capture_k2.py run [iters] [dtype]        # ONE K2 launch of the bench's C2 chain set (24 576 chains)
    python tools/capture_k2.py parse <ncu.csv> <out.json> [mix.json]   # -> profiles/r2_k2.json (read by bench.py: roofline.traffic)
    python tools/capture_k2.py mix <full.ncu-rep> <mix.json>           # dynamic packed-FP mix from the source page of a --set full capture

ncu's ``smsp__sass_thread_inst_executed_op_{ffma,fmul,fadd}_pred_on`` count the SCALAR instructions only: the packed
FFMA2 / FMUL2 / FADD2 this kernel is made of are not in them (measured: they give 6 % of the algorithmic flops).
Executed flops are therefore formed from the launch's warp-instruction count and the executed-instruction mix of the
source page of a ``--set full`` capture of the same kernel (``mix``): ``32 x inst_executed x sum(share x flops)``.

On the GPU box (one gpurun call; the plain run first, as the profiling recipe demands):

    python tools/capture_k2.py run 3000 > gpurun_out/k2_plain.log 2>&1 &&
    ncu --metrics $(python tools/capture_k2.py metrics) --clock-control none -k regex:k2_mcmc -c 1 --csv \\
        --log-file gpurun_out/k2_counters.csv python tools/capture_k2.py run 3000
    python tools/capture_k2.py parse gpurun_out/k2_counters.csv profiles/r2_k2.json

Only a dozen counters are collected (a few replays of the 0.4 s launch), not ``--set full`` (about 40 replays: the
round-1 attempt on the 3000-iteration launch ran into the call's time limit).  The JSON holds the DRAM bytes of the
launch and the executed FP32 instruction counters next to the algorithmic flops of the same launch (the ``run`` mode
prints them from the device work counters), so executed and algorithmic flops can be read side by side.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__inst_executed.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum",
    "smsp__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_lsu.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size",
]


def run(iters, dtype):
    import numpy as np
    import torch
    from vamp_b200 import batch, engine, flops, synth
    ctx = engine.default_context(0)
    reg = synth.config_c2(region_id=2)
    regs = engine.Regions(ctx, [(reg["nu"], reg["flux"], reg["noise"])])
    models = [(0, n) for n in range(1, 7)]
    per = 4 * 1024
    item_region = np.zeros(len(models) * per, np.int32)
    item_ncomp = np.repeat(np.array([n for _, n in models], np.int32), per)
    uid = batch.model_chain_uids(np.array([2]), models, 4, 1024)
    order = np.lexsort((np.arange(item_ncomp.size), item_ncomp))
    s = engine.Sampler(regs, item_region[order], item_ncomp[order], uid[order], voigt=True, sd_free=True, dtype=dtype, seed=1,
                       burn=iters // 10, thin=15, tune_interval=50, tune_throughout=True, cmax=6)
    s.init()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s.run(iters)                       # the one K2 launch
    e1.record()
    torch.cuda.synchronize()
    w = s.work()
    ms = e0.elapsed_time(e1)
    fl = flops.k2_flops(w, dtype, voigt=True)
    print(json.dumps({"iters": iters, "dtype": dtype, "chains": int(s.B), "kernel_ms_events": ms, "work": w,
                      "algorithmic_flops": fl, "algorithmic_tflops": fl / ms / 1e9,
                      "evals_per_sec": w["shape_pix"] / ms * 1e3}))


def mix(rep, out):
    """Executed-instruction mix of the kernel in a ``--set full --import-source on`` capture (no GPU needed)."""
    import collections
    import io
    import subprocess
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
    h = collections.Counter()
    for r in rows[2:]:
        op = r[isrc].strip().split()
        if not op:
            continue
        o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
        h[o] += int(r[iex])
    tot = float(sum(h.values()))
    res = {"source": os.path.basename(rep), "kernel": rows[0][1] if len(rows[0]) > 1 else None, "warp_instructions": tot,
           "share": {k: v / tot for k, v in h.most_common(40)}}
    packed = h["FFMA2"] + h["FMUL2"] + h["FADD2"]
    res["packed_share"] = packed / tot
    res["flops_per_packed_thread_inst"] = (4.0 * h["FFMA2"] + 2.0 * (h["FMUL2"] + h["FADD2"])) / max(packed, 1)
    with open(out, "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "share"}))


def parse(path, out, mixfile=None):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    im, iv, iu, ik = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Kernel Name")
    d, units, kernel = {}, {}, None
    for r in rows[start + 1:]:
        if len(r) <= iv:
            continue
        kernel = r[ik]
        d[r[im]] = float(r[iv].replace(",", "")) if r[iv] not in ("", "n/a") else None
        units[r[im]] = r[iu]
    scale = {"ns": 1e-9, "nsecond": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0}
    t = d.get("gpu__time_duration.sum", 0.0) * scale.get(units.get("gpu__time_duration.sum", "ns"), 1e-9)
    byte = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def b(k):
        return (d.get(k) or 0.0) * byte.get(units.get(k, "byte"), 1.0)

    ffma, fmul, fadd = (d.get("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % o) or 0.0 for o in ("ffma", "fmul", "fadd"))
    res = {"source": os.path.basename(path), "kernel": kernel, "gpu_time_s_under_ncu": t,
           "dram_bytes_read": b("dram__bytes_read.sum"), "dram_bytes_write": b("dram__bytes_write.sum"),
           "traffic_bytes": b("dram__bytes_read.sum") + b("dram__bytes_write.sum"),
           "thread_inst_ffma": ffma, "thread_inst_fmul": fmul, "thread_inst_fadd": fadd,
           "scalar_fp32_flops": 2.0 * ffma + fmul + fadd,         # ncu's roofline convention: scalar instructions only
           "metrics": d, "units": units}
    plain = os.path.join(os.path.dirname(path), "k2_plain.log")
    if os.path.exists(plain):
        for line in open(plain):
            if line.startswith("{"):
                res["plain_run"] = json.loads(line)
    if mixfile and os.path.exists(mixfile):
        # executed FP32 arithmetic: warp instructions of the launch x 32 lanes x the executed-instruction mix (source
        # page of the --set full capture of the same kernel) x flops per instruction (FFMA2 4, FMUL2 / FADD2 2,
        # FFMA 2, FMUL / FADD 1).  MUFU, compares, min / max and conversions are not arithmetic flops here.
        mx = json.load(open(mixfile))
        sh = mx["share"]
        w = {"FFMA2": 4.0, "FMUL2": 2.0, "FADD2": 2.0, "FFMA": 2.0, "FMUL": 1.0, "FADD": 1.0}
        per_inst = sum(w[k] * sh.get(k, 0.0) for k in w)
        res["mix"] = {"source": mx["source"], "share": {k: sh.get(k, 0.0) for k in list(w) + ["MUFU", "LDS", "STS", "SHFL", "BRA", "IMAD"]},
                      "fp32_flops_per_warp_lane_instruction": per_inst}
        res["executed_fp32_flops"] = 32.0 * (d.get("smsp__inst_executed.sum") or 0.0) * per_inst
        # consistency of the mix with the counters: scalar FFMA thread instructions either way
        res["scalar_ffma_from_mix"] = 32.0 * (d.get("smsp__inst_executed.sum") or 0.0) * sh.get("FFMA", 0.0)
    if res.get("plain_run"):
        res["algorithmic_flops"] = res["plain_run"]["algorithmic_flops"]
        if "executed_fp32_flops" in res:
            res["executed_over_algorithmic"] = res["executed_fp32_flops"] / res["algorithmic_flops"]
    with open(out, "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps({k: v for k, v in res.items() if k not in ("metrics", "units")}, indent=1))


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "run"
    if mode == "metrics":
        print(",".join(METRICS))
    elif mode == "parse":
        parse(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    elif mode == "mix":
        mix(sys.argv[2], sys.argv[3])
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 3000, sys.argv[3] if len(sys.argv) > 3 else "f32")
