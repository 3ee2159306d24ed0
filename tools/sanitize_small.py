#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck): every kernel on tiny inputs, including the multi-lane batch path."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgbd_visualodometry_b200 import orb
from rgbd_visualodometry_b200.synth import synth_frame, synth_descriptors, synth_map_queries
ctx = orb.Context(150, 1.2, 8, 200, 150, 130)
frames = [synth_frame(150, 200, 7000 + i) for i in range(130)]
k, d, n = ctx.detect_and_compute_batch(frames)                      # lanes
k1, d1 = ctx.detect_and_compute(synth_frame(149, 197, 3))           # per-level kernels, odd size
t = synth_descriptors(333, 1); q = synth_map_queries(t, 517, 2)
m = ctx.match(q, t); m2 = ctx.knn_match2(q, t)
cnt = np.array([len(t), 0, 97], np.int32)
tr = np.zeros((3, 400, 32), np.uint8); tr[0, :333] = t; tr[2, :97] = t[:97]
ms = ctx.match_sets(q, tr, cnt)
print("ok", int(n.sum()), len(k1), len(m), m2.shape, ms.shape)
