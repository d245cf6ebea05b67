#!/usr/bin/env python
"""Top warp-stall sites of one kernel from `ncu -i X.ncu-rep --page source --csv` (optionally gzipped).
   python tools/ncu_source_top.py FILE.csv[.gz] KERNEL_INDEX [N]      (CPU; kernel index = order in the file, SASS views only)"""
import csv, gzip, sys
path, kidx = sys.argv[1], int(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
op = gzip.open if path.endswith(".gz") else open
kernels, cur = [], None
with op(path, "rt", newline="") as f:
    for row in csv.reader(f):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None:
            cur["rows"].append(row)
sass = [k for k in kernels if k["hdr"] and k["hdr"][0] == "Address"]
print(f"{len(kernels)} views, {len(sass)} SASS views")
for i, k in enumerate(sass):
    tot = sum(int(r[k['hdr'].index('# Samples')] or 0) for r in k["rows"])
    print(f"  [{i}] {k['name'][:90]}  instr={len(k['rows'])} samples={tot}")
k = sass[kidx]
h = k["hdr"]
iS, iSrc = h.index("# Samples"), h.index("Source")
stall_cols = [(j, c) for j, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
rows = k["rows"]
tot = sum(int(r[iS] or 0) for r in rows)
print(f"\nkernel [{kidx}] {k['name'][:100]}\n total samples {tot}")
agg = {}
for r in rows:
    for j, c in stall_cols:
        agg[c] = agg.get(c, 0) + int(r[j] or 0)
print(" stall totals:", ", ".join(f"{c[6:]}={v} ({v/max(tot,1):.0%})" for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(rows)), key=lambda i: -int(rows[i][iS] or 0))[:topn]
for i in sorted(order):
    r = rows[i]
    st = sorted(((int(r[j] or 0), c[6:]) for j, c in stall_cols), reverse=True)[:2]
    print(f" {i:5d} {int(r[iS]):6d} {int(r[iS])/max(tot,1):6.1%}  {r[iSrc].strip()[:70]:70s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")
