#!/usr/bin/env python
"""Generate tests/golden/*.npz from the real OpenCV operators (cv2, the in-image build of the library
the reference calls at src/frontend.cpp:153 and :187).  Run in the build container; the fixtures travel.

  orb_<case>.npz  : input sha256, nfeatures, cv2 keypoints (28-byte cv::KeyPoint records), descriptors
  match_<case>.npz: input sha256s, cv2 BFMatcher(NORM_HAMMING).match and knnMatch(k=2) results
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import MATCH_CASES, ORB_CASES, sha  # noqa: E402
from oracle import oracle as O  # noqa: E402  (only for the cv2->record converters)

out = os.path.join(ROOT, "tests", "golden")
for name, (mk, n) in ORB_CASES.items():
    img = mk()
    k, d = cv2.ORB_create(n, 1.2, 8).detectAndCompute(img, None)
    d = d if d is not None else np.zeros((0, 32), np.uint8)
    np.savez_compressed(os.path.join(out, f"orb_{name}.npz"), img_sha=sha(img), nfeatures=n, cv2_version=cv2.__version__,
                        keypoints=O.cv2_keypoints_to_array(k), descriptors=d)
    print(name, img.shape, n, len(k))
bf = cv2.BFMatcher(cv2.NORM_HAMMING)
for name, (mq, mt) in MATCH_CASES.items():
    q, t = mq(), mt()
    m = O.cv2_matches_to_array(bf.match(q, t))
    knn = bf.knnMatch(q, t, k=2)
    k2 = np.zeros((len(knn), 2), O.MATCH_DTYPE)
    k2["trainIdx"] = -1
    for i, row in enumerate(knn):
        for j, mm in enumerate(row):
            k2[i, j] = (mm.queryIdx, mm.trainIdx, mm.imgIdx, mm.distance)
    np.savez_compressed(os.path.join(out, f"match_{name}.npz"), q_sha=sha(q), t_sha=sha(t), match=m, knn2=k2)
    print(name, q.shape, t.shape, len(m))
