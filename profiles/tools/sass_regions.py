#!/usr/bin/env python3
"""Break an `ncu --page source --csv` SASS profile down by kernel region, by
opcode class and by stall reason.

usage: sass_regions.py <ncu_source.csv> <nvdisasm -gi -c output> [kernel-substring]

nvdisasm -gi annotates each instruction with its inline chain; the outermost
frame (a line of resident.cuh inside the kernel body) tells which section of
the scheduler the instruction belongs to; out-of-line functions are their own
regions. Joins on the instruction offset.
"""
import csv
import re
import sys
from collections import defaultdict

csv.field_size_limit(10 ** 9)

# (first line, last line, name) of the kernel body in resident.cuh
REGIONS = []


def load_regions(spec):
    """`name:first_line,name:first_line,...` (lines of resident.cuh)."""
    marks = []
    for item in spec.split(","):
        n, l = item.split(":")
        marks.append((int(l), n))
    return sorted(marks)


def region_of(marks, line):
    name = "body<%d" % (marks[0][0] if marks else 0)
    for first, n in marks:
        if line >= first:
            name = n
    return name


def opclass(op):
    op = op.split()[0] if op else "?"
    if op.startswith("@"):
        op = op.split()[1] if len(op.split()) > 1 else op
    base = op.split(".")[0]
    if base in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX"):
        return "fp64"
    if base in ("MUFU",):
        return "mufu"
    if base in ("LDL", "STL"):
        return base
    if base in ("LDS", "STS", "LDSM", "ATOMS"):
        return "smem"
    if base in ("LDG", "LD", "STG", "ST", "ATOMG", "RED", "LDC", "LDCU", "ULDC"):
        return "gmem/const"
    if base in ("BRA", "BSSY", "BSYNC", "CALL", "RET", "EXIT", "WARPSYNC",
                "BREAK", "BMOV", "JMP", "BRX", "YIELD", "NANOSLEEP"):
        return "control"
    if base in ("BAR", "SYNCS", "MEMBAR", "FENCE", "ERRBAR"):
        return "sync"
    if base in ("IMAD", "IADD3", "IADD", "LOP3", "SHF", "LEA", "ISETP", "SEL",
                "IABS", "FLO", "POPC", "PRMT", "PLOP3", "P2R", "R2P", "VOTE",
                "SHFL", "VOTEU", "UIADD3", "ULOP3", "USHF", "UMOV", "UISETP",
                "UIMAD", "USEL", "ULEA", "R2UR", "S2R", "S2UR", "CS2R", "BREV",
                "IMNMX", "VIMNMX", "VIADD", "UFLO", "UPOPC", "REDUX", "MATCH",
                "LEPC", "UPLOP3", "UPRMT", "UP2UR", "UR2UP", "SGXT", "UIADD",
                "UVIMNMX", "UVIADD", "USGXT", "FSEL", "FSETP", "FMUL", "FADD",
                "FFMA", "I2F", "F2I", "F2F", "I2FP", "F2FP", "FMNMX", "FCHK"):
        return "int/fp32"
    if base in ("MOV", "NOP"):
        return base.lower()
    return "other:" + base


def parse_sass(path, kernel):
    """offset -> (region key, innermost (file, line), opcode text)."""
    table = {}
    inside = False
    chain = []
    func = None
    after_inst = False
    for raw in open(path, errors="replace"):
        if raw.startswith("//-----") and ".text." in raw:
            inside = True
            func = raw.split(".text.")[1].strip()
            chain = []
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', raw)
        if m:
            f = m.group(1).split("/")[-1]
            if after_inst:
                chain = []
                after_inst = False
            if m.group(3) is not None:
                chain.append((m.group(3).split("/")[-1], int(m.group(4))))
            else:
                chain.append((f, int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", raw)
        if m:
            off = int(m.group(1), 16)
            table[(func, off)] = (list(chain), m.group(2))
            after_inst = True
    return table


def main():
    src, sass = sys.argv[1], sys.argv[2]
    kernel = sys.argv[3] if len(sys.argv) > 3 else "transport_resident"
    marks = load_regions(sys.argv[4] if len(sys.argv) > 4 else
                         "prologue:574,del_warp:638,step_prologue:698,pool:713,"
                         "service:799,barrier:897,ehs:914,step:952,epilogue:996")
    table = parse_sass(sass, kernel)
    funcs = sorted({f for f, _ in table})
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    header = rows[h]
    col = {name: i for i, name in enumerate(header)}
    stall_cols = [c for c in header if c.startswith("stall_") and "Not Issued" not in c]

    # The source page lists the kernel first, then its out-of-line functions,
    # contiguous in memory in link order; match the kernel's section by size.
    kfunc = next(f for f in funcs if kernel in f)
    koffs = sorted(o for f, o in table if f == kfunc)
    base = None
    agg = defaultdict(lambda: defaultdict(float))
    ops = defaultdict(lambda: defaultdict(float))
    stall_by_op = defaultdict(lambda: defaultdict(float))
    total = defaultdict(float)
    for r in rows[h + 1:]:
        if len(r) < len(header):
            continue
        try:
            addr = int(r[col["Address"]], 16)
        except ValueError:
            continue
        if base is None:
            base = addr
        off = addr - base
        ent = table.get((kfunc, off))
        text = r[col["Source"]].strip()
        if ent is None or off > koffs[-1]:
            region = "out-of-line"
        else:
            chain, _ = ent
            # outermost frame = last element of the chain
            outer = chain[-1] if chain else ("?", 0)
            if outer[0] == "resident.cuh":
                region = region_of(marks, outer[1])
            else:
                region = "?" + outer[0]
        inst = float(r[col["Instructions Executed"]] or 0)
        thr = float(r[col["Thread Instructions Executed"]] or 0)
        smp = float(r[col["# Samples"]] or 0)
        a = agg[region]
        a["inst"] += inst
        a["thr"] += thr
        a["smp"] += smp
        for c in stall_cols:
            v = float(r[col[c]] or 0)
            a[c] += v
            total[c] += v
        total["inst"] += inst
        total["thr"] += thr
        total["smp"] += smp
        oc = opclass(text)
        ops[region][oc] += inst
        ops["ALL"][oc] += inst
        stall_by_op[oc]["smp"] += smp
        stall_by_op[oc]["long"] += float(r[col["stall_long_sb"]] or 0)
        stall_by_op[oc]["inst"] += inst

    print("region: inst%  samples%  active | top stalls (share of region samples)")
    for region, a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"]):
        st = sorted(((a[c], c) for c in stall_cols), reverse=True)[:4]
        print("  %-14s inst %5.1f%%  samples %5.1f%%  active %4.1f | %s" % (
            region, 100 * a["inst"] / total["inst"], 100 * a["smp"] / total["smp"],
            a["thr"] / max(a["inst"], 1),
            "  ".join("%s %.0f%%" % (c[6:], 100 * v / max(a["smp"], 1)) for v, c in st)))
        o = ops[region]
        tot = sum(o.values()) or 1
        print("      ops: " + "  ".join("%s %.0f%%" % (k, 100 * v / tot)
              for k, v in sorted(o.items(), key=lambda kv: -kv[1])[:9]))
    print("\nopcode classes (all regions): inst%  samples%  long_sb share")
    for oc, v in sorted(stall_by_op.items(), key=lambda kv: -kv[1]["smp"])[:16]:
        print("  %-14s inst %5.1f%%  samples %5.1f%%  long_sb %5.1f%%" % (
            oc, 100 * v["inst"] / total["inst"], 100 * v["smp"] / total["smp"],
            100 * v["long"] / max(total["stall_long_sb"], 1)))


if __name__ == "__main__":
    main()
