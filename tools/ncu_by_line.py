"""Attribute an ncu source-page csv (SASS view, tools/gpu_ncu_cmd.sh) to SOURCE lines: instruction offsets are matched
against `nvdisasm -g -c` of the object file (line info from -lineinfo).
usage: python tools/ncu_by_line.py <source.csv.gz> <object.o> <kernel-substring>  [top]"""
import collections
import csv
import gzip
import os
import re
import subprocess
import sys
import tempfile


def line_table(obj, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    table = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        on, cur = False, None
        for ln in txt.splitlines():
            if ln.startswith(".text."):
                on = kernel in ln
                continue
            if not on:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main(path, obj, kernel, top=30):
    table = line_table(obj, kernel)
    op = gzip.open if path.endswith(".gz") else open
    rows = list(csv.reader(op(path, "rt")))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    base = int(body[0][ci["Address"]], 16)
    ex, smp = collections.Counter(), collections.Counter()
    miss = 0
    for r in body:
        off = int(r[ci["Address"]], 16) - base
        src = table.get(off, (None, None))[0]
        if src is None:
            miss += 1
        ex[src] += int(r[ci["Instructions Executed"]])
        smp[src] += int(r[ci["# Samples"]])
    tot, tots = sum(ex.values()), max(1, sum(smp.values()))
    print("instructions executed %d, samples %d, unmatched SASS lines %d" % (tot, tots, miss))
    print("%-26s %12s %7s %8s %7s" % ("source line", "executed", "%", "samples", "%"))
    for src, v in ex.most_common(top):
        print("%-26s %12d %6.1f%% %8d %6.1f%%" % ("%s:%s" % src if src else "?", v, 100.0 * v / tot, smp[src], 100.0 * smp[src] / tots))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 30)
