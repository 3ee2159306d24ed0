#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of both families on tiny inputs,
the multi-lane host path with pageable buffers (staging threads), maps, and the matcher's kernels.
  compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgbd_visualodometry_b200 import orb
from rgbd_visualodometry_b200.synth import synth_frame, synth_descriptors, synth_map_queries
ctx = orb.Context(150, 1.2, 8, 200, 150, 130)
frames = [synth_frame(150, 200, 7000 + i) for i in range(130)]
k, d, n = ctx.detect_and_compute_batch(frames)                      # lanes
k1, d1 = ctx.detect_and_compute(synth_frame(149, 197, 3))           # CTA-cooperative family (small launch), odd size
ctx.force_kernels(0)
k0, d0 = ctx.detect_and_compute(synth_frame(149, 197, 3))           # warp-private TMA family on the same frame
assert k0.tobytes() == k1.tobytes() and np.array_equal(d0, d1)
kb, db, nb = ctx.detect_and_compute_batch(frames[:9])                # ... and on a batch (describe groups, side stream)
assert all(kb[i, :nb[i]].tobytes() == k[i, :n[i]].tobytes() for i in range(9))
ctx.force_kernels(-1)
kp, dp, cp, bp = ctx.extract_match_batch(frames[:40], [synth_descriptors(70, 9)], 200)   # host lanes + matches per lane
t = synth_descriptors(333, 1); q = synth_map_queries(t, 517, 2)
m = ctx.match(q, t); m2 = ctx.knn_match2(q, t)
cnt = np.array([len(t), 0, 97], np.int32)
tr = np.zeros((3, 400, 32), np.uint8); tr[0, :333] = t; tr[2, :97] = t[:97]
ms = ctx.match_sets(q, tr, cnt)
print("ok", int(n.sum()), len(k1), len(m), m2.shape, ms.shape)
