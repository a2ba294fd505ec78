"""Loop structure of a kernel from an ncu source-page csv (SASS view): contiguous runs of SASS lines with the same executed
count (= one loop body / straight-line region), their instruction and sample shares, and the hottest lines with their top
stall reasons.  usage: python tools/ncu_blocks.py <source.csv.gz> [min_instr] [top]"""
import csv
import gzip
import sys


def main(path, min_instr=150000, top=16):
    rows = list(csv.reader(gzip.open(path, "rt") if path.endswith(".gz") else open(path)))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    ex = [int(r[ci["Instructions Executed"]]) for r in body]
    sm = [int(r[ci["# Samples"]]) for r in body]
    src = [r[ci["Source"]].strip() for r in body]
    print("total instr", sum(ex), "samples", sum(sm), "lines", len(body))
    i = 0
    while i < len(body):
        j = i
        while j + 1 < len(body) and abs(ex[j + 1] - ex[i]) <= 0.02 * max(1, ex[i]):
            j += 1
        te, ts = sum(ex[i:j + 1]), sum(sm[i:j + 1])
        if te > min_instr or ts > 0.015 * sum(sm):
            print("lines %5d-%5d n=%4d exec/line %8d  total %9d (%4.1f%%) samples %5d (%4.1f%%)  first: %s" % (
                i, j, j - i + 1, ex[i], te, 100.0 * te / sum(ex), ts, 100.0 * ts / sum(sm), src[i][:40]))
        i = j + 1
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    order = sorted(range(len(body)), key=lambda k: -sm[k])[:top]
    for k in sorted(order):
        r = body[k]
        st = sorted(((int(r[ci[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
        print("%5d smp %4d exec %7d  %-52s %s" % (k, sm[k], ex[k], src[k][:52], st))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 150000, int(sys.argv[3]) if len(sys.argv) > 3 else 16)
