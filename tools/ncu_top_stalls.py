#!/usr/bin/env python
"""Top stalled SASS instructions of the first kernel in an ncu report (needs -lineinfo / --import-source)."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur); continue
    if cur is None: continue
    if r and r[0] == "Address": cur["hdr"] = r; continue
    cur["rows"].append(r)
b = blocks[int(sys.argv[3]) if len(sys.argv) > 3 else 0]
h = b["hdr"]; si, sm, ie = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
rs = [r for r in b["rows"] if len(r) > sm and r[sm].isdigit()]
tot = sum(int(r[sm]) for r in rs)
print(b["name"][:100], "total samples", tot)
cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
agg = {}
for r in rs:
    for i in cols:
        if r[i].isdigit(): agg[h[i][6:]] = agg.get(h[i][6:], 0) + int(r[i])
print("stall totals:", sorted(agg.items(), key=lambda x: -x[1])[:8])
for r in sorted(rs, key=lambda r: -int(r[sm]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    st = sorted(((h[i][6:], int(r[i])) for i in cols if r[i].isdigit() and int(r[i]) > 0), key=lambda x: -x[1])[:3]
    print(f"{int(r[sm]):6d} {100*int(r[sm])/tot:5.1f}% exec={r[ie]:>8s} {r[si].strip()[:72]:72s} {st}")
