#!/usr/bin/env python
"""Stall samples of one kernel of an .ncu-rep aggregated by SOURCE LINE (via nvdisasm line info of the current
liborbx.so -- the kernel's SASS must be unchanged since the capture; instruction counts are checked).
  python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_select [N]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "rgbd_visualodometry_b200", "liborbx.so")
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
blk = rows[hi[0] + 1: (hi[1] - 1 if len(hi) > 1 else len(rows))]
hdr = rows[hi[0]]
si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
data = [r for r in blk if len(r) > si and r[si].isdigit()]
with tempfile.TemporaryDirectory() as td:
    subprocess.check_call(["cuobjdump", "-xelf", "all", SO], cwd=td, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, cubin)], capture_output=True, text=True).stdout
lines, infn, cur = [], False, None
for ln in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        infn = kern in m.group(1); cur = None; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = int(m.group(2)); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln): lines.append(cur)
print("instructions: report", len(data), "binary", len(lines))
agg, ex = collections.Counter(), collections.Counter()
for i, r in enumerate(data):
    l = lines[i] if i < len(lines) else None
    agg[l] += int(r[si]); ex[l] += int(r[ie])
tot = sum(agg.values())
srcl = open(os.path.join(ROOT, "rgbd_visualodometry_b200", "csrc", "orbx_kernels.cuh")).read().splitlines()
for l, c in agg.most_common(n):
    text = srcl[l - 1].strip()[:90] if l and l <= len(srcl) else ""
    print(f"{str(l):>5s} {c:6d} {100.0*c/tot:5.1f}% exec={ex[l]:>9d}  {text}")
