#!/usr/bin/env python3
"""Per source line executed-instruction shares of one kernel, from
`ncu --page source --csv` joined with `nvdisasm -g -c` (see sass_lines.py).
usage: sass_hot.py <ncu_source.csv> <sass> <kernel-substring> <src-dir> [min-share%] [file ...]"""
import collections
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
csv.field_size_limit(10 ** 9)
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "sass_lines.py")).read()
ns = {}
exec(src.split("def main")[0], ns)
table = ns["parse_sass"](sys.argv[2], sys.argv[3])
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, ii, it = (hdr.index("Address"), hdr.index("Instructions Executed"),
              hdr.index("Thread Instructions Executed"))
isamp = hdr.index("# Samples")
base = None
inst = collections.Counter()
thr = collections.Counter()
smp = collections.Counter()
for r in rows[2:]:
    a = int(r[ia], 16)
    if base is None:
        base = a
    off = a - base
    if off in table:
        (f, l), _ = table[off]
        inst[(f, l)] += int(r[ii])
        thr[(f, l)] += int(r[it])
        smp[(f, l)] += int(r[isamp])
tot = sum(inst.values())
tots = sum(smp.values())
srcdir = sys.argv[4]
minshare = float(sys.argv[5]) if len(sys.argv) > 5 else 0.3
files = sys.argv[6:] or sorted({f for f, _ in inst})
for f in files:
    path = os.path.join(srcdir, f)
    lines = open(path).read().split("\n") if os.path.exists(path) else []
    sub = sum(c for (ff, _), c in inst.items() if ff == f)
    print("===== %s  inst %.1f%%" % (f, 100. * sub / tot))
    for (ff, l), c in sorted(inst.items(), key=lambda kv: kv[0][1]):
        if ff == f and 100. * c / tot >= minshare:
            text = lines[l - 1].strip()[:70] if l - 1 < len(lines) else ""
            print("%5d inst %5.2f%% smp %5.2f%% lanes %4.1f  %s" % (
                l, 100. * c / tot, 100. * smp[(ff, l)] / max(tots, 1),
                thr[(ff, l)] / max(c, 1), text))
