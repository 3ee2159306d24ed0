#!/usr/bin/env python
"""Summarise the source page of an .ncu-rep for one kernel: hottest SASS instructions by stall samples.
  python tools/ncu_src.py gpurun_out/prof.ncu-rep k_blur [N]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several launches may be concatenated: take the first block
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
blk = rows[hi[0] + 1: (hi[1] - 1 if len(hi) > 1 else len(rows))]
hdr = rows[hi[0]]
si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
data = [r for r in blk if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in data)
print("total samples", tot, "instructions", len(data), "warp-inst executed", sum(int(r[ie]) for r in data))
for i, r in sorted(enumerate(data), key=lambda t: -int(t[1][si]))[:n]:
    print(f"{i:5d} {int(r[si]):6d} {100.0*int(r[si])/tot:5.1f}%  exec={r[ie]:>9s}  {r[src].strip()[:100]}")
if "--loads" in sys.argv:
    for i, r in enumerate(data):
        if "LDG" in r[src] or "STG" in r[src] or "BAR" in r[src]:
            print(f"{i:5d} {int(r[si]):6d} {r[src].strip()[:90]}")
