"""ctypes wrapper over oracle/liborb_oracle.so (the CPU restatement in orb_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's cpu legs,
never by the product package (rgbd_visualodometry_b200/).  See the header of orb_oracle.c for what
it restates (cv::ORB::detectAndCompute, src/frontend.cpp:153; BFMatcher(NORM_HAMMING)::match, :187)
and how it is pinned (bit-exact vs the in-image cv2 4.13.0 and tests/golden/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liborb_oracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
MATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
CAND_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("response", "<f4")])
MAX_LEVELS = 16


class _Dump(C.Structure):
    _fields_ = [("levels", C.c_void_p * MAX_LEVELS), ("blurred", C.c_void_p * MAX_LEVELS),
                ("fast", C.c_void_p * MAX_LEVELS), ("fast_cap", C.c_int),
                ("n_fast", C.c_int * MAX_LEVELS), ("n_sel1", C.c_int * MAX_LEVELS), ("n_sel2", C.c_int * MAX_LEVELS)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "orb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liborb_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orbo_harris.restype = C.c_float
        _lib.orbo_ic_angle.restype = C.c_float
        _lib.orbo_fast_atan2.restype = C.c_float
        _lib.orbo_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib.orbo_sincosf.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        _lib.orbo_brief.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
        _lib.orbo_detect_and_compute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_float,
                                                 C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p]
        _lib.orbo_filter_matches.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def gray(bgr: np.ndarray) -> np.ndarray:
    bgr = np.ascontiguousarray(bgr)
    h, w = bgr.shape[:2]
    out = np.empty((h, w), np.uint8)
    lib().orbo_gray(_p(bgr), w, h, C.c_size_t(w * 3), _p(out))
    return out


def level_geometry(w: int, h: int, nlevels: int = 8, scale_factor: float = 1.2):
    ws = (C.c_int * nlevels)(); hs = (C.c_int * nlevels)(); sc = (C.c_float * nlevels)()
    lib().orbo_level_geometry(w, h, nlevels, C.c_float(scale_factor), ws, hs, sc)
    return list(ws), list(hs), np.array(list(sc), np.float32)


def quotas(nfeatures: int, nlevels: int = 8, scale_factor: float = 1.2):
    n = (C.c_int * nlevels)()
    lib().orbo_quotas(nfeatures, C.c_float(scale_factor), nlevels, n)
    return list(n)


def resize_exact(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    src = np.ascontiguousarray(src)
    out = np.empty((dh, dw), np.uint8)
    lib().orbo_resize_exact(_p(src), src.shape[1], src.shape[0], _p(out), dw, dh)
    return out


def fast_nms(img: np.ndarray, t: int = 20) -> np.ndarray:
    img = np.ascontiguousarray(img)
    h, w = img.shape
    cap = ((w + 1) // 2) * ((h + 1) // 2) + 1
    out = np.zeros(cap, CAND_DTYPE)
    n = lib().orbo_fast_nms(_p(img), w, h, t, _p(out), cap)
    return out[:n]


def fast_score_map(img: np.ndarray, t: int = 20) -> np.ndarray:
    img = np.ascontiguousarray(img)
    out = np.empty_like(img)
    lib().orbo_fast_score_map(_p(img), img.shape[1], img.shape[0], t, _p(out))
    return out


def retain_best(cands: np.ndarray, m: int) -> np.ndarray:
    v = np.ascontiguousarray(cands.copy())
    n = lib().orbo_retain_best(_p(v), len(v), m)
    if n < 0:
        raise RuntimeError("heap_select fallback flagged")
    return v[:n]


def harris(img: np.ndarray, x: int, y: int) -> float:
    return lib().orbo_harris(_p(img), img.shape[1], int(x), int(y))


def ic_angle(img: np.ndarray, x: int, y: int) -> float:
    return lib().orbo_ic_angle(_p(img), img.shape[1], int(x), int(y))


def blur7(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img)
    out = np.empty_like(img)
    lib().orbo_blur7(_p(img), img.shape[1], img.shape[0], _p(out))
    return out


def sincosf(ang: float):
    s = C.c_float(); c = C.c_float()
    lib().orbo_sincosf(C.c_float(ang), C.byref(s), C.byref(c))
    return s.value, c.value


def detect_and_compute(img: np.ndarray, nfeatures: int = 500, scale_factor: float = 1.2, nlevels: int = 8,
                       cap: int | None = None, dump: bool = False):
    """Returns (keypoints[KP_DTYPE], descriptors[n,32] u8) and, if dump, a dict of per-level stages."""
    img = np.ascontiguousarray(img)
    h, w = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    cap = cap if cap is not None else max(4 * nfeatures, 64)
    d = None
    keep = {}
    if dump:
        ws, hs, _ = level_geometry(w, h, nlevels, scale_factor)
        d = _Dump()
        d.fast_cap = 0
        keep = {"levels": [], "blurred": [], "fast": []}
        fcap = ((w + 1) // 2) * ((h + 1) // 2) + 1
        d.fast_cap = fcap
        for l in range(nlevels):
            lv = np.zeros((max(hs[l], 1), max(ws[l], 1)), np.uint8); bl = np.zeros_like(lv); fa = np.zeros(fcap, CAND_DTYPE)
            keep["levels"].append(lv); keep["blurred"].append(bl); keep["fast"].append(fa)
            d.levels[l] = lv.ctypes.data; d.blurred[l] = bl.ctypes.data; d.fast[l] = fa.ctypes.data
    while True:
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        rc = lib().orbo_detect_and_compute(_p(img), w, h, C.c_size_t(img.strides[0]), ch, nfeatures, C.c_float(scale_factor),
                                           nlevels, _p(kps), _p(desc), cap, C.byref(n), C.byref(d) if d is not None else None)
        if rc == -2:
            cap = n.value
            continue
        if rc != 0:
            raise RuntimeError(f"oracle detect_and_compute rc={rc}")
        break
    out = (kps[:n.value].copy(), desc[:n.value].copy())
    if dump:
        keep["n_fast"] = list(d.n_fast)[:nlevels]; keep["n_sel1"] = list(d.n_sel1)[:nlevels]; keep["n_sel2"] = list(d.n_sel2)[:nlevels]
        keep["fast"] = [keep["fast"][l][:keep["n_fast"][l]] for l in range(nlevels)]
        return out + (keep,)
    return out


def match_hamming(query: np.ndarray, train: np.ndarray) -> np.ndarray:
    query = np.ascontiguousarray(query, np.uint8); train = np.ascontiguousarray(train, np.uint8)
    nq, nt = len(query), len(train)
    out = np.zeros(max(nq, 1), MATCH_DTYPE)
    n = lib().orbo_match_hamming(_p(query), nq, _p(train), nt, _p(out))
    return out[:n]


def match_hamming_knn2(query: np.ndarray, train: np.ndarray) -> np.ndarray:
    query = np.ascontiguousarray(query, np.uint8); train = np.ascontiguousarray(train, np.uint8)
    nq, nt = len(query), len(train)
    out = np.zeros((max(nq, 1), 2), MATCH_DTYPE)
    n = lib().orbo_match_hamming_knn2(_p(query), nq, _p(train), nt, _p(out))
    return out[:n]


def filter_matches(matches: np.ndarray, ratio: float = 2.0) -> np.ndarray:
    matches = np.ascontiguousarray(matches)
    out = np.zeros(max(len(matches), 1), MATCH_DTYPE)
    n = lib().orbo_filter_matches(_p(matches), len(matches), C.c_float(ratio), _p(out))
    return out[:n]


# ---- helpers to compare against cv2 (the real OpenCV operators the reference calls) ----
def cv2_keypoints_to_array(kps) -> np.ndarray:
    a = np.zeros(len(kps), KP_DTYPE)
    for i, k in enumerate(kps):
        a[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id)
    return a


def cv2_matches_to_array(ms) -> np.ndarray:
    a = np.zeros(len(ms), MATCH_DTYPE)
    for i, m in enumerate(ms):
        a[i] = (m.queryIdx, m.trainIdx, m.imgIdx, m.distance)
    return a
