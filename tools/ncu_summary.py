"""Summarise an exported ncu raw-page csv (tools/gpu_ncu.sh) into the handful of numbers DESIGN.md / profiles/ cite."""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg ", "sm__cycles_elapsed.max", "smsp__inst_executed_op_local", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__inst_executed_pipe_lsu.sum", "lts__t_bytes.sum ", "l1tex__t_bytes.sum ", "smsp__average_warp", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__sass_thread_inst_executed_op_ffma", "sm__sass_thread_inst_executed_op_fp32", "derived__smsp__sass_thread_inst_executed_op",
        "smsp__pcsamp_warps_issue_stalled", "smsp__average_warps_issue_stalled", "sm__cycles_active.avg", "gpc__cycles_elapsed.avg.per_second", "sm__pipe_fma_cycles_active.avg.pct", "sm__inst_executed_pipe_fmaheavy", "sm__pipe_fmaheavy", "sm__inst_executed_pipe_fmalite","smsp__inst_executed_pipe_fma","smsp__inst_executed_pipe_xu", "smsp__inst_executed_pipe_alu", "sm__inst_executed_pipe"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("== kernel:", vals[hdr.index("Kernel Name")][:110] if "Kernel Name" in hdr else "?")
        for h, u, v in zip(hdr, units, vals):
            if any(w.strip() in h for w in WANT):
                print("  %-95s %-14s %s" % (h, u, v))


if __name__ == "__main__":
    main(sys.argv[1])
