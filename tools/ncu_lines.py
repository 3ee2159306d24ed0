#!/usr/bin/env python
"""Executed instructions and stall samples of one kernel of an .ncu-rep aggregated by SOURCE LINE (via nvdisasm line info of
the current liborbx.so -- the kernel's SASS must be unchanged since the capture; instruction counts are checked).
  python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_select [N] [--by-exec]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "rgbd_visualodometry_b200", "liborbx.so")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
rep, kern = args[0], args[1]
n = int(args[2]) if len(args) > 2 else 30
by_exec = "--by-exec" in sys.argv
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[hi[0]]
# several launches of the kernel may be in the report: take the one that executed the most instructions
blks = [rows[h + 1: (hi[j + 1] - 1 if j + 1 < len(hi) else len(rows))] for j, h in enumerate(hi)]
_ie = hdr.index("Instructions Executed")
blk = max(blks, key=lambda b: sum(int(r[_ie]) for r in b if len(r) > _ie and r[_ie].isdigit()))
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
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln): lines.append(cur)
print("instructions: report", len(data), "binary", len(lines))
agg, ex = collections.Counter(), collections.Counter()
for i, r in enumerate(data):
    l = lines[i] if i < len(lines) else None
    agg[l] += int(r[si]); ex[l] += int(r[ie])
tot, tex = sum(agg.values()), sum(ex.values())
cache = {}
def text(l):
    if not l: return ""
    f, k = l
    if f not in cache:
        p = os.path.join(ROOT, "rgbd_visualodometry_b200", "csrc", f)
        cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return cache[f][k - 1].strip()[:80] if k <= len(cache[f]) else ""
order = sorted(agg, key=lambda l: -(ex[l] if by_exec else agg[l]))[:n]
for l in order:
    print(f"{(l[0][5:9] + ':' + str(l[1])) if l else 'None':>10s} samples {agg[l]:6d} {100.0*agg[l]/tot:5.1f}%  exec {ex[l]/1e6:8.2f}M {100.0*ex[l]/tex:5.1f}%  {text(l)}")
