"""Summarise an .ncu-rep (one kernel, --set full) into the text kept under profiles/.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
    "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "smsp__inst_executed_op_global_ld.sum", "l1tex__t_sector_hit_rate.pct",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name"))
        for k in KEYS:
            if k in d:
                print("  %-72s %s %s" % (k, d[k], u[k]))
        stalls = [(float(v), h) for h, v in d.items()
                  if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and v not in ("", "n/a")]
        print("  warp stall reasons (avg warps stalled per issue-active cycle):")
        for v, h in sorted(stalls, reverse=True)[:8]:
            print("    %-60s %.3f" % (h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), v))
        sc = [float(d.get("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % o, 0) or 0) for o in ("ffma", "fmul", "fadd")]
        print("  scalar FP32 thread instructions (ffma, fmul, fadd): %.4g %.4g %.4g - the packed FFMA2 / FMUL2 / FADD2 are NOT in these "
              "counters; executed flops: tools/capture_k2.py (instruction mix x warp instructions)" % tuple(sc))


if __name__ == "__main__":
    main(sys.argv[1])
