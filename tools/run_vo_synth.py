#!/usr/bin/env python
"""BASELINE config 1 in miniature: the reference's tracking loop (FrontEnd::TrackingHandler, src/frontend.cpp:94-144) on a
synthetic TUM-fr1-shaped RGB-D sequence, with the hot path behind the orbx C-ABI and everything else on the host exactly as
in the reference (PnP = cv2.solvePnPRansac standing in for cv::solvePnPRansac, src/frontend.cpp:217-260).

Per frame: pose := previous pose; extract (asynchronously: the next frame is submitted before this one is processed);
MatchKeyPointsInTrackingMap + EstimatePosePnP twice (coarse / fine, as the reference does); key-frames every `kf` frames
create new map points from unmatched keypoints with depth (CreateNewMappoints).  `--frontend cv2` runs the identical loop
with OpenCV's own operators and a numpy visibility filter, for an end-to-end comparison of the two front-ends.

  python tools/run_vo_synth.py [--frames 40] [--frontend orbx|cv2|both]
"""
import argparse, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgbd_visualodometry_b200.synth import synth_frame, synth_depth

W, H = 640, 480
CAM = (517.3, 516.5, 318.6, 255.3)             # config/default.yaml:10-13
DEPTH_SCALE, NFEAT, RATIO, PLANE_Z = 5000.0, 500, 2.0, 2.0


def make_sequence(n, seed=0):
    """A camera sliding parallel to a fronto-parallel textured plane at 2 m: frame i sees the texture window at integer
    offset (ox, oy) px, i.e. the camera centre sits at (ox - ox0) * Z / fx, (oy - oy0) * Z / fy, 0 relative to frame 0."""
    big = synth_frame(H + 192, W + 192, seed)
    rng = np.random.default_rng(seed + 1)
    ox = oy = 96.0
    out = []
    first = None
    for i in range(n):
        ox = float(np.clip(ox + rng.uniform(-3, 3), 0, 192)); oy = float(np.clip(oy + rng.uniform(-3, 3), 0, 192))
        x0, y0 = int(round(ox)), int(round(oy))
        if first is None: first = (x0, y0)                   # the world frame is the first camera (its pose is the identity)
        centre = np.array([(x0 - first[0]) * PLANE_Z / CAM[0], (y0 - first[1]) * PLANE_Z / CAM[1], 0.0])
        out.append((np.ascontiguousarray(big[y0:y0 + H, x0:x0 + W]), synth_depth(H, W, seed + i, PLANE_Z, int(DEPTH_SCALE)), centre))
    return out


def pnp(pts3, pts2, pose):
    import cv2
    K = np.array([[CAM[0], 0, CAM[2]], [0, CAM[1], CAM[3]], [0, 0, 1.0]])
    if len(pts3) < 6:
        return pose, 0
    cv2.setRNGSeed(0)                                         # RANSAC draws: same sequence for both front-ends
    rvec, _ = cv2.Rodrigues(pose[:, :3]); tvec = pose[:, 3:4].copy()
    ok, rvec, tvec, inl = cv2.solvePnPRansac(pts3.astype(np.float64), pts2.astype(np.float64), K, None, rvec, tvec, True, 100, 4.0, 0.99)
    if not ok or inl is None:
        return pose, 0
    R, _ = cv2.Rodrigues(rvec)
    return np.concatenate([R, tvec.reshape(3, 1)], axis=1), len(inl)


class OrbxFrontEnd:
    def __init__(self):
        from rgbd_visualodometry_b200 import orb
        self.ctx = orb.Context(NFEAT, 1.2, 8, W, H, 1)
    def submit(self, img): self.ctx.submit_frame(img)
    def collect(self): return self.ctx.collect_frame()
    def add_points(self, ids, kp_index, pos, norm, kps, desc): self.ctx.map_upsert_from_frame(ids, 0, kp_index, pos, norm)
    def match(self, ids, pose, kps, desc):
        cand, m, _, _ = self.ctx.track_match(ids, pose, CAM, W, H, train=None, frame=0, match_ratio=RATIO)
        return cand, m
    def backproject(self, kps, depth, pose): return self.ctx.backproject(kps, depth, DEPTH_SCALE, CAM, pose)


class Cv2FrontEnd:
    """The reference's operators (cv2) + the numpy restatement of its glue -- the comparison arm."""
    def __init__(self):
        import cv2
        from oracle import oracle as O, track_oracle as T
        self.O, self.T = O, T
        self.orb = cv2.ORB_create(NFEAT, 1.2, 8); self.bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        self.q = []; self.desc, self.pos, self.norm = {}, {}, {}
    def submit(self, img): self.q.append(img)
    def collect(self):
        k, d = self.orb.detectAndCompute(self.q.pop(0), None)
        return self.O.cv2_keypoints_to_array(k), d
    def add_points(self, ids, kp_index, pos, norm, kps, desc):
        for i, j, p, nv in zip(ids, kp_index, pos, norm): self.desc[int(i)], self.pos[int(i)], self.norm[int(i)] = desc[j], p, nv
    def match(self, ids, pose, kps, desc):
        P = np.array([self.pos[int(i)] for i in ids]); N = np.array([self.norm[int(i)] for i in ids]); D = np.array([self.desc[int(i)] for i in ids])
        fn = lambda q, t: self.O.cv2_matches_to_array(self.bf.match(q, t))
        cand, m, _, _ = self.T.track_match(pose, CAM, W, H, P, N, np.zeros(len(ids), bool), D, desc, RATIO, fn)
        return cand, (m if m is not None else np.zeros(0, self.O.MATCH_DTYPE))
    def backproject(self, kps, depth, pose): return self.T.backproject(kps, depth, DEPTH_SCALE, CAM, pose)


def run(frontend, seq, kf=5):
    fe = OrbxFrontEnd() if frontend == "orbx" else Cv2FrontEnd()
    pose = np.concatenate([np.eye(3), np.zeros((3, 1))], axis=1)
    ids = np.zeros(0, np.int64); next_id = 0
    log = []
    fe.submit(seq[0][0])
    t0 = time.perf_counter()
    for i, (img, depth, centre) in enumerate(seq):
        if i + 1 < len(seq): fe.submit(seq[i + 1][0])         # the GPU extracts frame i+1 while the host tracks frame i
        kps, desc = fe.collect()
        matched = np.zeros(len(kps), bool)
        nm = ninl = 0
        if len(ids):
            for _ in range(2):                                # coarse, then fine (src/frontend.cpp:101-108)
                cand, m = fe.match(ids, pose, kps, desc)
                nm = len(m)
                pid = ids[cand[m["queryIdx"]]] if nm else np.zeros(0, np.int64)
                pts3 = np.array([run.pos[int(p)] for p in pid]).reshape(-1, 3)
                pts2 = np.stack([kps["x"][m["trainIdx"]], kps["y"][m["trainIdx"]]], axis=1) if nm else np.zeros((0, 2))
                pose, ninl = pnp(pts3, pts2, pose)
            matched[m["trainIdx"]] = True
        if i % kf == 0:                                       # key-frame: CreateNewMappoints (src/frontend.cpp:372-406)
            new = np.nonzero(~matched)[0].astype(np.int32)
            pos, valid = fe.backproject(kps[new], depth, pose)
            new, pos = new[valid], pos[valid]
            c = -(pose[:, :3].T @ pose[:, 3])
            nv = pos - c; nv /= np.linalg.norm(nv, axis=1, keepdims=True)      # Mappoint::AddObservedByKeyframe, mappoint.h:63
            nid = np.arange(next_id, next_id + len(new), dtype=np.int64); next_id += len(new)
            for a, p in zip(nid, pos): run.pos[int(a)] = p
            fe.add_points(nid, new, pos, nv, kps, desc)
            ids = np.concatenate([ids, nid])
        est = -(pose[:, :3].T @ pose[:, 3])
        log.append((i, len(kps), nm, ninl, est, centre))
    dt = time.perf_counter() - t0
    return log, dt
run.pos = {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--frontend", default="both", choices=["orbx", "cv2", "both"])
    a = ap.parse_args()
    seq = make_sequence(a.frames)
    res = {}
    for fe in (["orbx", "cv2"] if a.frontend == "both" else [a.frontend]):
        run.pos = {}
        log, dt = run(fe, seq)
        err = np.array([np.linalg.norm(e - c) for (_, _, _, _, e, c) in log])
        res[fe] = log
        print(f"{fe:5s}: {a.frames} frames in {dt*1e3:.0f} ms ({a.frames/dt:.1f} frames/s incl. host PnP)  matches/frame {np.mean([l[2] for l in log[1:]]):.0f}  "
              f"inliers {np.mean([l[3] for l in log[1:]]):.0f}  camera-centre error mean {err.mean()*1e3:.2f} mm  max {err.max()*1e3:.2f} mm")
    if len(res) == 2:
        same = all(a_[1:4] == b_[1:4] and np.allclose(a_[4], b_[4], atol=1e-9) for a_, b_ in zip(res["orbx"], res["cv2"]))
        print("front-ends agree frame by frame (keypoints, matches, inliers, pose):", same)


if __name__ == "__main__":
    main()
