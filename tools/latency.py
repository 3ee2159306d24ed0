#!/usr/bin/env python
"""Single-frame latency of the drop-in calls (BASELINE config 1 shape: 640x480, 500 features; one extraction + two
matches per frame, src/frontend.cpp:98-108), host buffers in and out, wall clock -- beside cv2 on one host core."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgbd_visualodometry_b200 import orb
from rgbd_visualodometry_b200.synth import synth_frame, synth_map_queries

frames = [synth_frame(480, 640, 100 + i) for i in range(16)]
ctx = orb.Context(500, 1.2, 8, 640, 480, 1)
k0, d0 = ctx.detect_and_compute(frames[0])
mp = synth_map_queries(d0, 1500, 3)


def timeit(fn, n=200):
    for _ in range(10): fn()
    t0 = time.perf_counter()
    for i in range(n): fn(i)
    return (time.perf_counter() - t0) / n * 1e3


state = {}
def ext(i=0): state["kd"] = ctx.detect_and_compute(frames[i % 16])
def mat(i=0): ctx.match(mp, state["kd"][1])
def both(i=0):
    ext(i); mat(i); mat(i)
t_e, t_m, t_b = timeit(ext), timeit(mat), timeit(both)
print(f"orbx  : extract {t_e:.3f} ms   match(1500 x {len(state['kd'][1])}) {t_m:.3f} ms   extract + 2 matches {t_b:.3f} ms / frame")
try:
    import cv2
    cv2.setNumThreads(1)
    o = cv2.ORB_create(500, 1.2, 8); bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    def cext(i=0): state["c"] = o.detectAndCompute(frames[i % 16], None)
    def cmat(i=0): bf.match(mp, state["c"][1])
    def cboth(i=0):
        cext(i); cmat(i); cmat(i)
    c_e, c_m, c_b = timeit(cext, 30), timeit(cmat, 30), timeit(cboth, 30)
    print(f"cv2   : extract {c_e:.3f} ms   match {c_m:.3f} ms   extract + 2 matches {c_b:.3f} ms / frame  (1 thread, cv2 {cv2.__version__})")
except Exception as e:
    print("cv2 unavailable:", e)
