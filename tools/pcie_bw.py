#!/usr/bin/env python
"""Pinned host <-> device copy bandwidth of the box (context for the e2e number: frames are uploaded every step).

  python tools/pcie_bw.py                                                       # one GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/pcie_bw.py
      # every rank copies at the same time: the aggregate is what the host's memory system / PCIe root complexes sustain,
      # i.e. the ceiling of the N-GPU end-to-end number (each rank uploads 236 MB per step)
Prints one JSON line (rank 0)."""
import json
import os
import time

import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 236 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
REPS = 20


def barrier():
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn):
    for _ in range(2):
        fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(REPS):
        fn()
    s2.synchronize()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                 # the slowest rank bounds the aggregate
    return REPS * n / float(t.item()) / 1e9                       # GB/s per rank (at the slowest rank's pace)


def both():
    d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d, non_blocking=True)


n_up, n_dn = 236060672, 36440080                                  # what one bench step moves (256 VGA frames up; records + matches down)


def step_mix():
    d[:n_up].copy_(h[:n_up], non_blocking=True)
    with torch.cuda.stream(s2):
        h2[:n_dn].copy_(d[:n_dn], non_blocking=True)


res = {"h2d": timed(lambda: d.copy_(h, non_blocking=True)), "d2h": timed(lambda: h.copy_(d, non_blocking=True)), "both_each_direction": timed(both)}
mix_steps_per_s = timed(step_mix) * 1e9 / n                        # timed() returns REPS * n bytes / seconds: back to steps per second
if rank == 0:
    print(json.dumps({"tool": "pcie_bw", "n_gpus": world, "bytes_per_copy": n, "host_cpus": os.cpu_count(),
                      "per_rank_gbs": res, "aggregate_gbs": {k: v * world for k, v in res.items()},
                      "bench_step_mix": {"h2d_bytes": n_up, "d2h_bytes": n_dn, "steps_per_s_per_rank": mix_steps_per_s,
                                         "e2e_frames_per_s_ceiling_640x480": mix_steps_per_s * 256 * world},
                      "note": "pinned memory, all ranks copy concurrently, wall clock of the slowest rank"}))
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
