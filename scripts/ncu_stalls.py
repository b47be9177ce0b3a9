#!/usr/bin/env python
"""Top warp-stall sites of the first kernel in an `ncu --page source --csv` dump, each with the SASS
just before it:  ncu_stalls.py src.csv [sites]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 8
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
end = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] in ("Kernel Name", "Address")), len(rows))
body = [r for r in rows[hi + 1:end] if len(r) >= len(hdr) - 2]
ss, ie, src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ss]) for r in body)
print("warp samples", tot)
for i in sorted(range(len(body)), key=lambda i: -int(body[i][ss]))[:top]:
    why = sorted(((int(body[i][c] or 0), h[6:]) for c, h in stall_cols), reverse=True)[:2]
    print(f"==== {int(body[i][ss]) / tot:.3f} of samples at SASS #{i}: " + ", ".join(f"{h} {v}" for v, h in why))
    for j in range(max(0, i - 4), i + 1):
        print(f"   {j:5d} exec {int(body[j][ie]):9d} samples {int(body[j][ss]):6d}  {body[j][src].strip()[:96]}")
