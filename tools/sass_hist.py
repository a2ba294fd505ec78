"""Opcode histogram of the hot loop from an exported ncu source page (SASS view): which instructions the time goes to."""
import csv, gzip, sys, collections
path = sys.argv[1]
rows = list(csv.reader(gzip.open(path, "rt")))
hdr = rows[1]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
body = rows[2:]
mx = max(int(r[iE]) for r in body)
hist, samp = collections.Counter(), collections.Counter()
tot = 0
for r in body:
    e = int(r[iE])
    op = r[iS].split()[0] if not r[iS].strip().startswith("@") else r[iS].split()[1]
    op = ".".join(op.split(".")[:2])
    hist[op] += e
    samp[op] += int(r[iSamp])
    tot += e
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 1
print("total warp-instr %d; per warp %.0f; static SASS lines %d; lines in hot loop (>=50%% of max exec) %d" % (
    tot, tot / nw, len(body), sum(1 for r in body if int(r[iE]) >= mx // 2)))
ts = sum(samp.values())
for op, c in hist.most_common(28):
    print("  %-18s %6.2f%% of instr   %6.2f%% of samples" % (op, 100.0 * c / tot, 100.0 * samp[op] / ts))
