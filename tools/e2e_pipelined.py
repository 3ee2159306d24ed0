#!/usr/bin/env python
"""End-to-end throughput of the host-buffer call with TWO contexts on two host threads taking alternate batches (the standard
double-buffering a caller with a stream of batches uses): the upload of batch i+1 runs under the kernels / download of batch i
ACROSS calls, which a single synchronous call cannot do for its own first upload and last download.
  python tools/e2e_pipelined.py [B] [steps] [pageable]"""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rgbd_visualodometry_b200 import orb  # noqa: E402
from rgbd_visualodometry_b200.synth import synth_descriptors  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
pageable = len(sys.argv) > 3 and sys.argv[3] == "pageable"
NCTX = int(os.environ.get("NCTX", "2"))
frames = bench.make_frames(B, 0)
qmap = synth_descriptors(bench.MAP_M, 3)


class Worker:
    def __init__(self):
        self.ctx = orb.Context(bench.NFEAT, bench.SCALE, bench.NLEVELS, bench.W, bench.H, B)
        pin = (lambda a: a) if pageable else (lambda a: a.pin_memory())
        self.inp = pin(torch.from_numpy(frames.copy()))
        self.kps = pin(torch.zeros((B, bench.CAP, 7), dtype=torch.float32))
        self.desc = pin(torch.zeros((B, bench.CAP, 32), dtype=torch.uint8))
        self.cnt = np.zeros(B, np.int32)
        self.map = torch.from_numpy(qmap).pin_memory()
        self.best = [pin(torch.zeros((B, bench.MAP_M, 4), dtype=torch.int32)) for _ in range(2)]
        self.ptrs = (C.c_void_p * B)(*[self.inp[i].data_ptr() for i in range(B)])
        self.qptrs = (C.c_void_p * 2)(self.map.data_ptr(), self.map.data_ptr())
        self.nqs = (C.c_int * 2)(bench.MAP_M, bench.MAP_M)
        self.bptrs = (C.c_void_p * 2)(*[t.data_ptr() for t in self.best])

    def call(self):
        rc = self.ctx.lib.orbx_extract_match_batch(self.ctx.h, self.ptrs, B, bench.W, bench.H, bench.W * 3, 3, self.kps.data_ptr(), self.desc.data_ptr(), bench.CAP,
                                                   self.cnt.ctypes.data, self.qptrs, self.nqs, 2, self.bptrs)
        assert rc == 0, self.ctx.lib.orbx_last_error(self.ctx.h)


ws = [Worker() for _ in range(NCTX)]
for w in ws:
    w.call(); w.call()


def run(n_each):
    def loop(w):
        for _ in range(n_each):
            w.call()
    th = [threading.Thread(target=loop, args=(w,)) for w in ws]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    return time.perf_counter() - t0


t1 = None
w0 = ws[0]
t0 = time.perf_counter()
for _ in range(steps):
    w0.call()
t1 = time.perf_counter() - t0
tn = run(steps)
print(f"B={B} {'pageable' if pageable else 'pinned'}: one context {B * steps / t1:.0f} frames/s; {NCTX} contexts on {NCTX} threads {NCTX * B * steps / tn:.0f} frames/s",
      "env=" + ",".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("ORBX_")))
