#!/usr/bin/env python
"""Minimal driver for ncu: N device-resident extraction (+ match) steps on B synthetic VGA frames."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rgbd_visualodometry_b200 import orb  # noqa: E402
from rgbd_visualodometry_b200.synth import synth_descriptors  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
frames = bench.make_frames(B, 0)
ctx = orb.Context(bench.NFEAT, bench.SCALE, bench.NLEVELS, bench.W, bench.H, B)
d_in = torch.from_numpy(frames).cuda()
d_k = torch.zeros((B, bench.CAP, 7), dtype=torch.float32, device="cuda")
d_d = torch.zeros((B, bench.CAP, 32), dtype=torch.uint8, device="cuda")
d_n = torch.zeros(B, dtype=torch.int32, device="cuda")
d_map = torch.from_numpy(synth_descriptors(bench.MAP_M, 3)).cuda()
d_best = torch.zeros((B, bench.MAP_M, 4), dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
for _ in range(steps):
    ctx.detect_and_compute_device(d_in.data_ptr(), B, bench.W, bench.H, bench.W * 3, bench.H * bench.W * 3, 3, d_k.data_ptr(), d_d.data_ptr(), bench.CAP, d_n.data_ptr())
    ctx.match_device_ragged(d_map.data_ptr(), bench.MAP_M, d_d.data_ptr(), bench.CAP, d_n.data_ptr(), B, d_best.data_ptr())
ctx.synchronize()
print("ok", d_n.cpu().numpy()[:4])
