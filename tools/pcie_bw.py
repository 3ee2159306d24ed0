#!/usr/bin/env python
"""Pinned host <-> device copy bandwidth of the box (context for the e2e number: frames are uploaded every step)."""
import torch
n = 236 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {10 * n / (e0.elapsed_time(e1) / 1e3) / 1e9:.1f} GB/s")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d, non_blocking=True)
e1.record(); s2.synchronize(); torch.cuda.synchronize()
print(f"H2D + D2H concurrently: {10 * n / (e0.elapsed_time(e1) / 1e3) / 1e9:.1f} GB/s each direction (approx)")
