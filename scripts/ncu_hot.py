#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: executed-instruction histogram by SASS opcode and the
hottest contiguous SASS regions.  usage: ncu_hot.py src.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# first kernel only
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
end = next((i for i in range(hdr_i + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
body = [r for r in rows[hdr_i + 1:end] if len(r) >= len(hdr) - 2]
ie, ss, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
tot = sum(int(r[ie]) for r in body)
tots = sum(int(r[ss]) for r in body)
print("instructions", len(body), "executed", tot, "samples", tots)
ops = collections.Counter()
for r in body:
    op = r[src].split()[0] if not r[src].strip().startswith("@") else r[src].split()[1]
    ops[op.split(".")[0]] += int(r[ie])
for op, c in ops.most_common(25):
    print(f"  {op:12s} {c:10d} {c/tot:.3f}")
print("--- hottest instructions")
for i, r in sorted(enumerate(body), key=lambda t: -int(t[1][ie]))[:top]:
    print(f"{i:5d} {int(r[ie]):9d} {int(r[ss]):6d}  {r[src].strip()[:90]}")
