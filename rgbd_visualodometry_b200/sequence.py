"""Sequence-level multi-GPU driver (the loop being sharded is app/run_vo.cpp:89-117): one ORDERED sequence of frames
is partitioned over the ranks with `shard_indices`, every rank pushes its frames through the host-buffer operator in batches,
and the per-frame results are gathered on every rank IN FRAME ORDER.  No data-path collective: frames are independent and the
map is replicated; torch.distributed only carries the gather of the results and the max-over-ranks of the timing.

The per-batch operator is passed in (`process_batch(frames) -> per-frame (kps, desc, matches)`), so the same driver runs on
`liborbx.so` (tools/run_sequence.py, bench) and -- in the world-size-2 gloo test on CPU -- on the oracle.
"""
from __future__ import annotations

import time
import zlib

import numpy as np

from .sharding import gather_in_frame_order, max_over_ranks, shard_indices


def frame_digest(kps: np.ndarray, desc: np.ndarray, matches: np.ndarray | None):
    """(count, crc32 of the keypoint records, of the descriptor bytes, of the match records): what is gathered for every
    frame of a long sequence (the full records of 10 k 4K frames would be gigabytes)."""
    m = zlib.crc32(np.ascontiguousarray(matches).tobytes()) if matches is not None else 0
    return (int(len(kps)), zlib.crc32(np.ascontiguousarray(kps).tobytes()), zlib.crc32(np.ascontiguousarray(desc).tobytes()), m)


def run_sharded_sequence(n_frames: int, make_frame, process_batch, rank: int, world: int, mode: str = "contiguous",
                         batch: int = 64, keep=(), group=None, device=None):
    """Process frames [0, n_frames) sharded over `world` ranks.

    make_frame(i) -> the i-th frame (numpy, host); process_batch(list of frames) -> list of (kps, desc, matches-or-None).
    Returns (digests, kept, seconds, op_seconds): `digests[i]` for EVERY frame in frame order (identical on all ranks), `kept[i]` =
    the full records of the frames listed in `keep` (gathered too), the wall time of the slowest rank's whole loop (frame
    production + operator + digests) and the part of it spent inside `process_batch` (max over ranks as well)."""
    mine = shard_indices(n_frames, rank, world, mode)
    keep = set(keep)
    local = []
    op_seconds = 0.0
    t0 = time.perf_counter()
    for s in range(0, len(mine), batch):
        chunk = mine[s:s + batch]
        frames = [make_frame(i) for i in chunk]
        t1 = time.perf_counter()
        out = process_batch(frames)
        op_seconds += time.perf_counter() - t1
        if len(out) != len(chunk):
            raise RuntimeError("process_batch returned a different number of frames")
        for i, (k, d, m) in zip(chunk, out):
            full = (np.ascontiguousarray(k).tobytes(), np.ascontiguousarray(d).tobytes(), np.ascontiguousarray(m).tobytes() if m is not None else b"") if i in keep else None
            local.append((frame_digest(k, d, m), full))
    seconds = time.perf_counter() - t0
    if world > 1:
        allres = gather_in_frame_order(mine, local, n_frames, group)
        seconds = max_over_ranks(seconds, device, group)
        op_seconds = max_over_ranks(op_seconds, device, group)
    else:
        allres = local
    digests = [r[0] for r in allres]
    kept = {i: allres[i][1] for i in keep if 0 <= i < n_frames}
    return digests, kept, seconds, op_seconds
