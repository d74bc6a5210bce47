#!/usr/bin/env python
"""Aggregate warp-stall reasons and the hottest SASS instructions of the first kernel in an .ncu-rep.
   python profiles/stalls.py gpurun_out/X.ncu-rep [top_n]"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
hdr, rows = None, []
for r in src:
    if r and r[0] == "Address":
        if hdr is not None: break
        hdr = r
    elif hdr is not None and r and r[0].startswith("0x"):
        rows.append(r)
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = Counter()
for r in rows:
    for i in stall_cols: tot[hdr[i]] += int(r[i] or 0)
S = sum(tot.values())
print("## stall reasons (all samples)"); 
for k, v in tot.most_common(): print(f"  {k:26s} {v:8d}  {100*v/S:5.1f} %")
i_s, i_n, i_e = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
print(f"\n## top {top} instructions by samples (idx, samples, executed, sass, main stall)")
order = sorted(range(len(rows)), key=lambda k: -int(rows[k][i_n]))[:top]
for k in sorted(order):
    r = rows[k]
    main = max(stall_cols, key=lambda i: int(r[i] or 0))
    print(f"  {k:5d} {int(r[i_n]):6d} {int(r[i_e]):9d}  {r[i_s].strip()[:70]:70s} {hdr[main]}")
