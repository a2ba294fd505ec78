"""Summarise an ncu source-page csv (tools/gpu_ncu_cmd.sh): stall samples by opcode and the hottest SASS lines;
also the share of samples inside the main loop (largest executed-count plateau)."""
import collections
import csv
import gzip
import sys


def main(path, top=25):
    op = gzip.open if path.endswith(".gz") else open
    rows = list(csv.reader(op(path, "rt")))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    tot = sum(int(r[ci["# Samples"]]) for r in body)
    byop, cnt = collections.Counter(), collections.Counter()
    stalls = [h for h in hdr if h.startswith("stall_")]
    bystall = collections.Counter()
    for r in body:
        toks = r[ci["Source"]].split()
        o = toks[1] if toks[0].startswith("@") else toks[0]
        o = o.split(".")[0]
        byop[o] += int(r[ci["# Samples"]])
        cnt[o] += int(r[ci["Instructions Executed"]])
        for s in stalls:
            bystall[s] += int(r[ci[s]] or 0)
    print("total samples", tot, " instructions executed", sum(cnt.values()))
    print("stall reasons:", ", ".join("%s %.1f%%" % (s[6:], 100.0 * v / max(1, sum(bystall.values()))) for s, v in bystall.most_common(10)))
    print("%-10s %9s %7s %12s %8s" % ("opcode", "samples", "%", "executed", "smp/exec"))
    for o, v in byop.most_common(18):
        print("%-10s %9d %6.1f%% %12d %8.3f" % (o, v, 100.0 * v / tot, cnt[o], v / max(1, cnt[o]) * 450))
    print("hottest lines:")
    order = sorted(range(len(body)), key=lambda i: -int(body[i][ci["# Samples"]]))[:top]
    for i in sorted(order):
        r = body[i]
        prev = body[i - 1][ci["Source"]].strip() if i else ""
        print("  %6d  %-60s <- %s" % (int(r[ci["# Samples"]]), r[ci["Source"]].strip()[:60], prev[:50]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
