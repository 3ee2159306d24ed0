"""N > 1 host logic on CPU: world-size-2 gloo group, frames sharded with no data-path collective, results gathered in
frame order, timing reduced as max over ranks.  The per-rank worker here is the CPU oracle (test stand-in for the
GPU operator, which needs a device); the driver under test is the one tools/run_sequence.py runs on the GPUs
(rgbd_visualodometry_b200/sequence.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rgbd_visualodometry_b200.sharding import shard_indices


def test_shard_indices_partition():
    for n in (0, 1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            for mode in ("contiguous", "round_robin"):
                parts = [shard_indices(n, r, world, mode) for r in range(world)]
                flat = sorted(i for p in parts for i in p)
                assert flat == list(range(n))
                assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert shard_indices(10, 1, 4) == [3, 4, 5]
    with pytest.raises(ValueError):
        shard_indices(4, 4, 4)


def _worker(rank, world, port, n_frames, mode, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from rgbd_visualodometry_b200.sharding import gather_in_frame_order, max_over_ranks, shard_indices
    from rgbd_visualodometry_b200.synth import synth_frame
    mine = shard_indices(n_frames, rank, world, mode)
    res = []
    for i in mine:
        k, d = O.detect_and_compute(synth_frame(120, 160, 900 + i), 100)
        res.append((k.tobytes(), d.tobytes()))
    allres = gather_in_frame_order(mine, res, n_frames)
    tmax = max_over_ranks(1.0 + rank)
    if rank == 0:
        q.put((allres, tmax))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["contiguous", "round_robin"])
def test_two_rank_gloo_gather(oracle, mode):
    from rgbd_visualodometry_b200.synth import synth_frame
    n_frames, world = 5, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    allres, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0                                   # max over ranks of (1 + rank)
    for i in range(n_frames):
        k, d = oracle.detect_and_compute(synth_frame(120, 160, 900 + i), 100)
        assert allres[i] == (k.tobytes(), d.tobytes()), f"frame {i} out of order or wrong"


def _seq_worker(rank, world, port, n_frames, mode, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from rgbd_visualodometry_b200.sequence import run_sharded_sequence
    from rgbd_visualodometry_b200.synth import synth_descriptors, synth_frame
    qmap = synth_descriptors(64, 3)

    def process(frames):
        out = []
        for fr in frames:
            k, d = O.detect_and_compute(fr, 80)
            out.append((k, d, O.match_hamming(qmap, d)))
        return out
    digests, kept, secs, op_secs = run_sharded_sequence(n_frames, lambda i: synth_frame(120, 160, 500 + i), process, rank, world, mode, batch=2, keep=(0, 3, n_frames - 1))
    assert 0 < op_secs <= secs
    q.put((rank, digests, kept, secs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["contiguous", "round_robin"])
def test_sharded_sequence_driver_two_ranks(oracle, mode):
    """The sequence driver: every rank ends up with the same frame-ordered digests, they equal the single-process result,
    the kept frames carry their full records, and the reported time is one number (the max over ranks)."""
    from rgbd_visualodometry_b200.sequence import frame_digest
    from rgbd_visualodometry_b200.synth import synth_descriptors, synth_frame
    n_frames, world = 7, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_seq_worker, args=(r, world, port, n_frames, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, d0, k0, s0), (_, d1, k1, s1) = got
    assert d0 == d1 and k0 == k1 and s0 == s1
    qmap = synth_descriptors(64, 3)
    for i in range(n_frames):
        k, d = oracle.detect_and_compute(synth_frame(120, 160, 500 + i), 80)
        m = oracle.match_hamming(qmap, d)
        assert d0[i] == frame_digest(k, d, m), f"frame {i} out of order or wrong"
        if i in (0, 3, n_frames - 1):
            assert k0[i] == (k.tobytes(), d.tobytes(), m.tobytes())
