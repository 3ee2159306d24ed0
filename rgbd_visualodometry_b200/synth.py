"""Seeded synthetic frames / descriptor sets shaped like the reference's inputs.

The reference reads TUM RGB-D PNGs with cv::imread (app/run_vo.cpp:91-92): 8UC3 BGR colour and
16UC1 depth (scale 5000, config/default.yaml:15).  No dataset ships with it and there is no
network, so every test and benchmark uses these generators (numpy only -- no OpenCV here).
"""
from __future__ import annotations

import numpy as np


def _box_blur3(a: np.ndarray, r: int) -> np.ndarray:
    """three box passes ~ gaussian; a is float32 HxWxC"""
    for _ in range(3):
        for ax in (0, 1):
            c = np.cumsum(np.pad(a, [(r + 1, r) if i == ax else (0, 0) for i in range(a.ndim)], mode="reflect"), axis=ax, dtype=np.float64)
            n = a.shape[ax]
            hi = np.take(c, np.arange(2 * r + 1, 2 * r + 1 + n), axis=ax)
            lo = np.take(c, np.arange(0, n), axis=ax)
            a = ((hi - lo) / (2 * r + 1)).astype(np.float32)
    return a


def synth_frame(h: int, w: int, seed: int, channels: int = 3, shapes: int | None = None) -> np.ndarray:
    """Textured BGR (or gray) frame: smooth noise + filled rectangles/discs + pixel noise.

    Gives ~1000 FAST corners at level 0 of a 640x480 frame (comparable to SURVEY.md Appendix C)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(h, w, channels), dtype=np.uint8).astype(np.float32)
    base = _box_blur3(base, 2)
    lo, hi = base.min(), base.max()
    img = (base - lo) * (255.0 / max(hi - lo, 1e-6))
    if shapes is None:
        shapes = max(8, int(200 * (h * w) / (480 * 640)))
    for _ in range(shapes):
        s = int(rng.integers(8, 60))
        cx, cy = int(rng.integers(0, w)), int(rng.integers(0, h))
        col = rng.integers(0, 256, size=channels).astype(np.float32)
        if rng.random() < 0.6:
            hh = int(s * rng.uniform(0.5, 1.5))
            y0, y1, x0, x1 = max(cy, 0), min(cy + hh, h), max(cx, 0), min(cx + s, w)
            img[y0:y1, x0:x1] = col
        else:
            r = s // 2
            y0, y1, x0, x1 = max(cy - r, 0), min(cy + r + 1, h), max(cx - r, 0), min(cx + r + 1, w)
            m = (np.arange(y0, y1)[:, None] - cy) ** 2 + (np.arange(x0, x1)[None, :] - cx) ** 2 <= r * r
            img[y0:y1, x0:x1][m] = col
    img = img + rng.integers(-6, 7, size=img.shape).astype(np.float32)
    out = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out if channels > 1 else out[:, :, 0]


def synth_depth(h: int, w: int, seed: int, metres: float = 2.0, scale: int = 5000) -> np.ndarray:
    """16UC1 fronto-parallel plane with 2% zero holes (TUM convention)."""
    rng = np.random.default_rng(seed + 7919)
    d = np.full((h, w), int(round(metres * scale)), dtype=np.uint16)
    d[rng.random((h, w)) < 0.02] = 0
    return d


def synth_sequence(h: int, w: int, n: int, seed: int = 0):
    """TUM-fr1-shaped sequence: a large texture seen through a slowly moving window (<=3 px/frame)."""
    big = synth_frame(h + 2 * 96, w + 2 * 96, seed)
    rng = np.random.default_rng(seed + 1)
    ox, oy = 96.0, 96.0
    for i in range(n):
        ox = float(np.clip(ox + rng.uniform(-3, 3), 0, 2 * 96))
        oy = float(np.clip(oy + rng.uniform(-3, 3), 0, 2 * 96))
        x0, y0 = int(round(ox)), int(round(oy))
        yield np.ascontiguousarray(big[y0:y0 + h, x0:x0 + w]), synth_depth(h, w, seed + i)


def synth_descriptors(n: int, seed: int) -> np.ndarray:
    """n x 32 uint8 i.i.d. uniform (Hamming distances ~ Binomial(256, 1/2))."""
    return np.random.default_rng(seed).integers(0, 256, size=(n, 32), dtype=np.uint8)


def synth_map_queries(train: np.ndarray, m: int, seed: int, flip: float = 0.10, noise_frac: float = 0.30) -> np.ndarray:
    """'realistic' map descriptors: noisy copies of train rows (true matches, many exact ties) + 30% pure noise."""
    rng = np.random.default_rng(seed)
    src = train[rng.integers(0, train.shape[0], size=m)]
    bits = np.unpackbits(src, axis=1)
    bits ^= (rng.random(bits.shape) < flip).astype(np.uint8)
    q = np.packbits(bits, axis=1)
    noise = rng.random(m) < noise_frac
    q[noise] = rng.integers(0, 256, size=(int(noise.sum()), 32), dtype=np.uint8)
    return np.ascontiguousarray(q)
