"""CPU restatement (numpy, float64) of the front-end glue either side of the matcher -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path
(rgbd_visualodometry_b200/) never does.

PARITY UNPINNED for these functions: they restate code of the reference itself (not of OpenCV), the reference cannot be
built in this container (Eigen3, Sophus, g2o, OpenCV C++ are absent -- SURVEY.md 8c) and its tests hold no vectors for
them, so there is nothing to pin the restatement against except the source text it follows, cited per function.
The reference evaluates these expressions in double precision (Eigen / Sophus) in a binary built with
-O3 -march=native, i.e. with FMA contraction at the compiler's discretion; parity is therefore defined to 1e-9
relative on the geometry and exact on the integer / index results away from decision boundaries.
"""
from __future__ import annotations

import numpy as np


def could_observe(T_c_w: np.ndarray, cam, cols: int, rows: int, pos: np.ndarray, norm: np.ndarray) -> np.ndarray:
    """Frame::IsCouldObserveMappoint, src/frame.cpp:70-91, for M points (pos, norm: [M, 3]); T_c_w: 3 x 4 [R | t]."""
    T = np.asarray(T_c_w, np.float64).reshape(3, 4)
    R, t = T[:, :3], T[:, 3]
    fx, fy, cx, cy = [float(v) for v in cam]
    pos = np.asarray(pos, np.float64).reshape(-1, 3)
    norm = np.asarray(norm, np.float64).reshape(-1, 3)
    out = np.zeros(len(pos), bool)
    center = -(R.T @ t)                                       # Frame::GetCamCenter: T_c_w.inverse().translation(), frame.h:54-56
    for i, (p, nv) in enumerate(zip(pos, norm)):
        pc = R @ p + t                                        # Camera::World2Camera, camera.cpp:39-42
        if pc[2] < 0:                                         # frame.cpp:73
            continue
        with np.errstate(divide="ignore", invalid="ignore"):
            u = fx * pc[0] / pc[2] + cx                       # Camera::Camera2Pixel, camera.cpp:49-55
            v = fy * pc[1] / pc[2] + cy
        if u < 0 or u >= cols or v < 0 or v >= rows:          # frame.cpp:78-81 (NaN compares false, as in C++)
            continue
        d = p - center                                        # frame.cpp:83-85
        with np.errstate(divide="ignore", invalid="ignore"):
            d = d / np.sqrt(d @ d)
            angle = np.arccos(d @ nv)
        if angle > np.pi / 6:                                 # frame.cpp:86 (NaN is not "> pi/6": the point is kept)
            continue
        out[i] = True
    return out


def track_match(T_c_w, cam, cols, rows, pos, norm, outlier, desc, train, match_ratio, match_fn):
    """FrontEnd::MatchKeyPointsInTrackingMap, src/frontend.cpp:169-211, for a tracking map given as arrays in iteration order.
    match_fn(query, train) -> DMatch records (the exact BFMatcher the north-star names).
    Returns (cand indices, kept matches, min_dis, max_dis)."""
    vis = could_observe(T_c_w, cam, cols, rows, pos, norm)
    cand = np.nonzero(vis & ~np.asarray(outlier, bool))[0].astype(np.int32)          # frontend.cpp:173-184
    if len(cand) == 0 or len(train) == 0:
        return cand, None, None, None
    m = match_fn(np.ascontiguousarray(desc[cand]), train)                            # frontend.cpp:187
    min_dis = np.float32(m["distance"].min())                                        # frontend.cpp:190-195
    max_dis = np.float32(max(np.float32(min_dis * np.float32(match_ratio)), np.float32(30.0)))   # frontend.cpp:196
    return cand, m[m["distance"] <= max_dis], float(min_dis), float(max_dis)         # frontend.cpp:203-211


def backproject(kps, depth, depth_scale, cam, T_c_w):
    """Frame::GetDepth (src/frame.cpp:43-67) + Camera::Pixel2World (src/camera.cpp:56-86) for keypoint records with fields
    x, y (float32).  Returns (pos_w [n, 3], valid [n]).  Lookups that would leave the image count as "no depth"."""
    T = np.asarray(T_c_w, np.float64).reshape(3, 4)
    R, t = T[:, :3], T[:, 3]
    fx, fy, cx, cy = [float(v) for v in cam]
    h, w = depth.shape
    n = len(kps)
    pos = np.zeros((n, 3), np.float64)
    valid = np.zeros(n, bool)
    for i in range(n):
        px, py = np.float32(kps["x"][i]), np.float32(kps["y"][i])
        x, y = int(np.rint(px)), int(np.rint(py))             # cvRound: round half to even
        d = 0
        if 0 <= x < w and 0 <= y < h:
            d = int(depth[y, x])
            if d == 0:
                for dx, dy in ((-1, 0), (0, -1), (1, 0), (0, 1)):       # frame.cpp:55-56
                    xx, yy = x + dx, y + dy
                    if 0 <= xx < w and 0 <= yy < h and depth[yy, xx] != 0:
                        d = int(depth[yy, xx])
                        break
        if d == 0:
            continue
        z = float(d) / float(np.float32(depth_scale))         # frame.cpp:50
        pc = np.array([(float(px) - cx) * z / fx, (float(py) - cy) * z / fy, z])    # camera.cpp:56-63
        pos[i] = R.T @ pc + (-(R.T @ t))                      # Camera2World: T_c_w.inverse() * p_c, camera.cpp:44-47
        valid[i] = True
    return pos, valid
