"""Frame sharding across GPUs: frames are independent units, the map is replicated, and there is
no data-path collective.  torch.distributed is used only to gather per-frame results on rank 0 in frame order and
to reduce timings (max over ranks)."""
from __future__ import annotations


def shard_indices(n_frames: int, rank: int, world: int, mode: str = "contiguous") -> list[int]:
    """Frame indices owned by `rank`.  'contiguous': near-equal chunks in frame order; 'round_robin': i % world."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if mode == "round_robin":
        return list(range(rank, n_frames, world))
    if mode != "contiguous":
        raise ValueError(mode)
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def gather_in_frame_order(local_indices, local_results, n_frames: int, group=None):
    """all_gather the (index, result) pairs and return the results as a frame-ordered list (every rank gets it)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    bucket = [None] * world
    dist.all_gather_object(bucket, list(zip(local_indices, local_results)), group=group)
    out = [None] * n_frames
    for part in bucket:
        for i, r in part:
            if out[i] is not None:
                raise RuntimeError(f"frame {i} produced by two ranks")
            out[i] = r
    if any(r is None for r in out):
        raise RuntimeError("some frames were produced by no rank")
    return out


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Multi-GPU numbers are timed on the device and reported as the max over ranks."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
