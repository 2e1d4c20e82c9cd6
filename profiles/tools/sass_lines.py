#!/usr/bin/env python3
"""Attribute an `ncu --page source --csv` SASS profile to source lines.

usage: sass_lines.py <ncu_source.csv> <nvdisasm -g -c output> [kernel-substring] [top]

ncu's CSV source page is per SASS instruction; nvdisasm -g annotates each
instruction group with `//## File "...", line N [inlined at ...]`. This joins
the two on the instruction offset and prints the hottest source lines by
warp-stall samples and by executed instructions, with the average number of
active threads (divergence) per line.
"""
import csv
import re
import sys
from collections import defaultdict

csv.field_size_limit(10 ** 9)


def parse_sass(path, kernel):
    """offset -> (file, line) for the first section whose name has `kernel`."""
    table = {}
    current = None
    inside = False
    for raw in open(path, errors="replace"):
        if raw.startswith("//-----") and ".text." in raw:
            inside = kernel in raw
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', raw)
        if m:
            current = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", raw)
        if m and current is not None:
            table[int(m.group(1), 16)] = (current, m.group(2))
    return table


def main():
    src, sass = sys.argv[1], sys.argv[2]
    kernel = sys.argv[3] if len(sys.argv) > 3 else "transport_kernel"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    table = parse_sass(sass, kernel)
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    header = rows[h]
    col = {name: i for i, name in enumerate(header)}
    agg = defaultdict(lambda: [0, 0, 0])  # samples, inst, thread inst
    by_file = defaultdict(lambda: [0, 0, 0])
    total = [0, 0, 0]
    base = None
    for r in rows[h + 1:]:
        if len(r) < len(header):
            continue
        try:
            addr = int(r[col["Address"]], 16)
        except ValueError:
            continue
        if base is None:
            base = addr
        key, _ = table.get(addr - base, (("?", 0), ""))
        vals = [int(float(r[col[c]] or 0)) for c in (
            "# Samples", "Instructions Executed",
            "Thread Instructions Executed")]
        for k in range(3):
            agg[key][k] += vals[k]
            by_file[key[0]][k] += vals[k]
            total[k] += vals[k]
    print("total samples %d, warp inst %d, thread inst %d, avg active threads "
          "%.1f" % (total[0], total[1], total[2], total[2] / max(total[1], 1)))
    print("\nby file:")
    for f, v in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
        print("  %-22s samples %5.1f%%  inst %5.1f%%  active %.1f" % (
            f, 100. * v[0] / total[0], 100. * v[1] / total[1],
            v[2] / max(v[1], 1)))
    print("\ntop lines by stall samples:")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("  %-22s:%-5d samples %5.2f%%  inst %5.2f%%  active %.1f" % (
            key[0], key[1], 100. * v[0] / total[0], 100. * v[1] / total[1],
            v[2] / max(v[1], 1)))


if __name__ == "__main__":
    main()
