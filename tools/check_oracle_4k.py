#!/usr/bin/env python
"""Pins the oracle to cv2 at BASELINE config 4/5 sizes (too slow for the default CPU suite): run manually."""
import os, sys
import numpy as np, cv2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from rgbd_visualodometry_b200.synth import synth_frame
for (h, w, n, seed) in ((1080, 1920, 2000, 5), (2160, 3840, 5000, 6)):
    img = synth_frame(h, w, seed, shapes=400)
    k, d = cv2.ORB_create(n, 1.2, 8).detectAndCompute(img, None)
    ko, do = O.detect_and_compute(img, n)
    ok = ko.tobytes() == O.cv2_keypoints_to_array(k).tobytes() and np.array_equal(d, do)
    print(f"{w}x{h} n={n}: cv2 {len(k)} oracle {len(ko)} -> {'EXACT' if ok else 'MISMATCH'}")
    assert ok
