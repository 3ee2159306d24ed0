#!/usr/bin/env python
"""Perf experiment: time the batched matcher with one pipeline role's work disabled (ORBX_MATCH_DBG bits:
1 = no epilogue TMEM loads, 2 = no MMA issue, 4 = no expansion).  Results are wrong in those modes by design."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgbd_visualodometry_b200 import orb
from rgbd_visualodometry_b200.synth import synth_descriptors
B, M, N, CAP = 256, 2048, 1000, 1280
ctx = orb.Context(1, 1.2, 1, 64, 64, 1)
dq = torch.from_numpy(synth_descriptors(M, 1)).cuda()
dt = torch.from_numpy(synth_descriptors(B * CAP, 2).reshape(B, CAP, 32)).cuda()
cnt = torch.full((B,), N, dtype=torch.int32, device="cuda")
best = torch.zeros((B, M, 4), dtype=torch.int32, device="cuda")
st = torch.cuda.ExternalStream(ctx.stream)
torch.cuda.synchronize()
for _ in range(5):
    ctx.match_device_ragged(dq.data_ptr(), M, dt.data_ptr(), CAP, cnt.data_ptr(), B, best.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(50):
    ctx.match_device_ragged(dq.data_ptr(), M, dt.data_ptr(), CAP, cnt.data_ptr(), B, best.data_ptr())
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
ops = 2.0 * 256 * M * N * B
print(f"dbg={os.environ.get('ORBX_MATCH_DBG','0')} {ms*1e3:.1f} us  {ops/ms/1e9:.0f} TOP/s  {ops/ms/1e9/4500*100:.1f}% of 4.5 POP/s")
if os.environ.get("ORBX_MATCH_TRACE"):
    tr = np.zeros(256, np.int64)
    ctx.lib.orbx_debug_match_trace(ctx.h, tr.ctypes.data)
    tr = tr.reshape(16, 16)
    base = tr[0, 0]
    names = ["iss:start", "iss:tempty", "iss:full", "iss:committed", "epi:start", "epi:tfull", "epi:ld_done", "epi:arrived", "exp:start", "exp:empty", "exp:arrived", "epi:computed", "epi:set_end", "epi:stored"]
    print("tile " + " ".join(f"{n:>13s}" for n in names))
    for i in range(16):
        print(f"{40+i:4d} " + " ".join(f"{(tr[i, k] - base) if tr[i, k] else -1:13d}" for k in range(14)))
