#!/usr/bin/env python
"""Contiguous SASS regions of one kernel with their share of executed instructions, from an
`ncu --page source --csv` dump: ncu_regions.py src.csv [min_share]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]; ie = hdr.index('Instructions Executed'); src = hdr.index('Source'); ss = hdr.index('# Samples')
end = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
body = [r for r in rows[hi + 1:end] if len(r) >= len(hdr) - 2]
tot = sum(int(r[ie]) for r in body); tots = sum(int(r[ss]) for r in body)
regs = []; start = 0
for i in range(1, len(body) + 1):
    if i == len(body) or abs(int(body[i][ie]) - int(body[i - 1][ie])) > 0.15 * max(int(body[i][ie]), int(body[i - 1][ie]), 1):
        regs.append((start, i)); start = i
print("instructions executed", tot, "samples", tots)
for a, b in regs:
    n = b - a; c = sum(int(r[ie]) for r in body[a:b]); s = sum(int(r[ss]) for r in body[a:b])
    if c / tot >= min_share:
        ops = {}
        for r in body[a:b]:
            t = r[src].split()
            o = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
            ops[o] = ops.get(o, 0) + 1
        print(f"[{a:4d},{b:4d}) n={n:3d} exec/instr={c // n:10d} share={c / tot:.3f} samples={s / max(tots, 1):.3f}",
              " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda t: -t[1])[:9]))
