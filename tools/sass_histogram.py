#!/usr/bin/env python
"""SASS opcode histogram of liborbx.so per kernel (CPU only: cuobjdump + nvdisasm): total instructions and the opcodes that
prove which hardware path a kernel uses (UTMALDG = TMA tile load, UTCIMMA / LDTM / STTM = tcgen05 MMA / tensor-memory
load / store, LDGSTS = cp.async, VIMNMX3 / VABSDIFF4 = byte / halfword SIMD, FFMA2 = packed FP32 ...).
  python tools/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "rgbd_visualodometry_b200", "liborbx.so")
KEYS = ["UTMALDG", "UTMASTG", "UBLKCP", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "LDGSTS", "SYNCS", "REDUX", "IDP", "VABSDIFF4", "VIMNMX3", "VIMNMX",
        "FFMA2", "FMUL2", "FADD2", "DFMA", "BAR", "ELECT", "R2UR"]
with tempfile.TemporaryDirectory() as td:
    subprocess.check_call(["cuobjdump", "-xelf", "all", SO], cwd=td, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", os.path.join(td, cubin)], capture_output=True, text=True).stdout
fn, per = None, collections.defaultdict(collections.Counter)
for ln in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and fn:
        per[fn][m.group(1).split(".")[0]] += 1


def demangle(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    return re.sub(r"\(.*", "", r)[:60]


print("# SASS opcode histogram of rgbd_visualodometry_b200/liborbx.so (sm_100a), per kernel: total instructions, then the opcodes that show")
print("# which hardware path a kernel uses.  Made by tools/sass_histogram.py (cuobjdump -xelf all + nvdisasm, counted per .text section).\n")
tot = collections.Counter()
for f, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values())):
    n = sum(c.values())
    if n < 50:
        continue
    print(f"{demangle(f):60s} {n:6d}  " + "  ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
    for k in KEYS:
        tot[k] += c[k]
print("\nlibrary total: " + "  ".join(f"{k}={tot[k]}" for k in KEYS if tot[k]))
