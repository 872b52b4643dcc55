#!/usr/bin/env python
"""Dynamic SASS opcode mix of the epilogue warps from an ncu source-page export.
usage: scripts/ncu_opmix.py gpurun_out/prof_TAG [n_top]"""
import csv, re, collections, sys
base = sys.argv[1]
rows = list(csv.reader(open(base + "_src.csv")))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
src, ex = ci["Source"], ci["Instructions Executed"]
epi = [i for i, r in enumerate(data) if "USETMAXREG.TRY_ALLOC" in r[src]][0]
tot = collections.Counter()
per_item = collections.Counter()
for r in data[epi:]:
    s = re.sub(r'^@!?U?P\d+\s+', '', r[src].strip())
    op = (s.split()[0] if s else '?').split('.')[0]
    tot[op] += int(r[ex])
    per_item[int(r[ex])] += 1
items = max((k for k, v in per_item.items() if v > 50), default=1)      # executions of a once-per-item line
total = sum(tot.values())
print("epilogue warp-instructions %d = %.1f per (warp, item)  [items x warps = %d]" % (total, total / items, items))
for op, n in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print("%-10s %12d %6.1f%%  %7.1f" % (op, n, 100 * n / total, n / items))
