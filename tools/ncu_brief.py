"""One-screen summary of an exported ncu raw page (+ optional source page): duration, launch shape, instruction totals,
IPC, issue / tensor-pipe activity, DRAM bytes, stall reasons per issue, opcode histogram of the executed instructions."""
import csv
import gzip
import sys
import collections

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]


def main(raw, source=None):
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
    print("kernel:", d.get("Kernel Name", "?"))
    for k in KEYS:
        if k in d:
            print("  %-82s %-10s %s" % (k, u[k], d[k]))
    st = [(h, float(d[h])) for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    print("  stall cycles per issued instruction:", ", ".join("%s %.2f" % (
        h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
        for h, v in sorted(st, key=lambda x: -x[1]) if v >= 0.05))
    if source:
        op = gzip.open if source.endswith(".gz") else open
        body = list(csv.reader(op(source, "rt")))
        h2 = body[1]
        ci = {h: i for i, h in enumerate(h2)}
        cnt = collections.Counter()
        for r in body[2:]:
            toks = r[ci["Source"]].split()
            o = toks[1] if toks[0].startswith("@") else toks[0]
            cnt[o.split(".")[0] + (".1688.TF32" if o.startswith("HMMA.1688.F32.TF32") else "")] += int(r[ci["Instructions Executed"]])
        tot = sum(cnt.values())
        print("  executed warp instructions by opcode (%d total): %s" % (tot, ", ".join("%s %.1f%%" % (o, 100.0 * v / tot) for o, v in cnt.most_common(14))))


if __name__ == "__main__":
    main(*sys.argv[1:3])
