"""Join an `ncu --page source --csv` SASS listing with `nvdisasm --print-line-info` of the same cubin
and aggregate executed warp-instructions per source line.

usage: sass_by_line.py <ncu_sass.csv> <nvdisasm.txt> <mangled-kernel-substring> <divisor> [min]
  divisor = number of units (e.g. events) the launch processed -> prints instructions per unit.
"""
import csv
import re
import sys
from collections import defaultdict

ncu_csv, dis, kname, div = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
minv = float(sys.argv[5]) if len(sys.argv) > 5 else 2.0

# nvdisasm: instruction offset -> innermost "file:line" (+ inlined-at chain collapsed to the outer line)
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l and l.endswith(":"))
loc = {}
cur = None
outer = None
for l in lines[start + 1:]:
    if l.startswith("//-----") or (l.startswith(".text.") and l.endswith(":")):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        f = m.group(1).split("/")[-1]
        cur = "%s:%s" % (f, m.group(2))
        chain = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        # outermost frame = the line of the kernel body this instruction belongs to
        outer = "%s:%s" % (chain[-1][0].split("/")[-1], chain[-1][1]) if chain else cur
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        loc[int(m.group(1), 16)] = (cur, outer, m.group(2).strip())

rows = list(csv.reader(open(ncu_csv)))
hdr = rows[1]
ia, ii, isrc = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Source")
ist = hdr.index("Warp Stall Sampling (All Samples)")
base = int(rows[2][ia], 16)
per_outer = defaultdict(float)
per_line = defaultdict(float)
per_line_stall = defaultdict(float)
per_op = defaultdict(float)
tot = 0.0
tot_st = 0.0
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break                        # only the first launch in the file
    if len(r) <= ii:
        continue
    off = int(r[ia], 16) - base
    n = float(r[ii])
    st = float(r[ist] or 0)
    c, o, txt = loc.get(off, ("?", None, r[isrc]))
    per_line[c] += n
    per_outer[o or c] += n
    per_line_stall[c] += st
    per_op[txt.split()[0] if not txt.startswith("@") else txt.split()[1]] += n
    tot += n
    tot_st += st
print("total warp-instructions / unit: %.1f" % (tot / div))
print("--- by source line (>= %.1f per unit) ---   instr/unit   stall-sample share" % minv)
for k, v in sorted(per_line.items(), key=lambda kv: -kv[1]):
    if v / div >= minv:
        print("%-32s %8.1f   %5.1f%%" % (k, v / div, 100 * per_line_stall[k] / max(tot_st, 1)))
print("--- by kernel-body line (inlined callees folded into their call site; use nvdisasm --print-line-info-inline) ---")
for k, v in sorted(per_outer.items(), key=lambda kv: kv[0]):
    if v / div >= minv / 2:
        print("%-32s %8.1f" % (k, v / div))
print("--- by opcode ---")
for k, v in sorted(per_op.items(), key=lambda kv: -kv[1])[:25]:
    print("%-24s %8.1f" % (k, v / div))
