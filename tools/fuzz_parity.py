#!/usr/bin/env python
"""Randomised GPU-vs-oracle parity sweep (run on the B200 box): many seeded synthetic frames of random sizes, feature
counts, textures (incl. rolled / flipped / noisy / blocky variants that provoke ties and the heap-select fallback).
Prints one line per mismatch and a summary; exit code 1 on any mismatch.
  python tools/fuzz_parity.py [n_frames] [seed]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from rgbd_visualodometry_b200 import orb
from rgbd_visualodometry_b200.synth import synth_frame, synth_map_queries

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
t0 = time.time()
ctxs = {}
for it in range(n_frames):
    h = int(rng.integers(90, 700)); w = int(rng.integers(100, 900))
    nf = int(rng.choice([100, 300, 500, 1000, 2000]))
    ch = int(rng.choice([1, 3]))
    img = synth_frame(h, w, int(rng.integers(0, 1 << 30)), ch)
    kind = int(rng.integers(0, 6))
    if kind == 1: img = np.ascontiguousarray(np.roll(img, (int(rng.integers(1, h)), int(rng.integers(1, w))), axis=(0, 1)))
    elif kind == 2: img = np.ascontiguousarray(img[:, ::-1])
    elif kind == 3: img = (img // 32 * 32).astype(np.uint8)                       # posterised: many exact ties
    elif kind == 4: img = np.clip(img.astype(np.int16) + rng.integers(-20, 21, img.shape), 0, 255).astype(np.uint8)
    elif kind == 5: img = np.ascontiguousarray(np.kron(img[: h // 2 + 1, : w // 2 + 1], np.ones((2, 2) + (() if img.ndim == 2 else (1,)), np.uint8))[:h, :w])
    key = nf
    if key not in ctxs: ctxs[key] = orb.Context(nf, 1.2, 8, 900, 700, 1)
    fam = it & 1                                                                  # both pyramid / FAST kernel families (orbx_debug_force_kernels)
    ctxs[key].force_kernels(fam)
    k, d = ctxs[key].detect_and_compute(img)
    ko, do = O.detect_and_compute(img, nf)
    ok = len(k) == len(ko) and k.tobytes() == ko.tobytes() and np.array_equal(d, do)
    if ok and len(k) > 0:
        q = synth_map_queries(d, int(rng.integers(1, 700)), int(rng.integers(0, 1 << 30)))
        m = ctxs[key].match(q, d)
        ok = m.tobytes() == O.match_hamming(q, do).tobytes()
    if not ok:
        bad += 1
        print(f"MISMATCH it={it} {w}x{h} ch={ch} nf={nf} kind={kind} kernels={fam} gpu={len(k)} oracle={len(ko)}", flush=True)
print(f"fuzz: {n_frames} frames (random sizes 100x90 .. 900x700, 1 / 3 channels, 100 .. 2000 features, 6 texture kinds, both kernel families alternating) "
      f"+ one match each vs the C oracle: {bad} mismatches, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
