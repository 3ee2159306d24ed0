#!/usr/bin/env python
"""One ordered synthetic sequence sharded over the GPUs of a box (BASELINE configs 4 / 5):

  python tools/run_sequence.py --config c4 --frames 512                          # 1 GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/run_sequence.py --config c5 --frames 10000

Frame i of the sequence goes to rank `shard_indices(...)` says; every rank runs `orbx_extract_match_batch` (host frames in;
keypoints, descriptors and the map's matches out) on its frames in batches; digests of every frame's results are gathered in
frame order, and rank 0 re-extracts a sample of frames ALONE (the 1-GPU result) and compares them bit for bit with what the
owning ranks produced.  Prints one JSON line.  Frames cycle through `--distinct` generated images (a 10 k-frame 4K sequence
does not fit host memory otherwise), rolled by a frame-dependent offset so that no two frames of the sequence are equal."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4", choices=["c1", "c2", "c4", "c5"])
    ap.add_argument("--frames", type=int, default=512)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--distinct", type=int, default=8)
    ap.add_argument("--mode", default="round_robin", choices=["contiguous", "round_robin"])
    ap.add_argument("--check", type=int, default=6, help="frames re-extracted on rank 0 alone and compared bit for bit")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from rgbd_visualodometry_b200 import orb
    from rgbd_visualodometry_b200.sequence import frame_digest, run_sharded_sequence
    from rgbd_visualodometry_b200.synth import synth_descriptors, synth_frame

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = bench.CONFIGS[args.config]
    w, h, nf = cfg["w"], cfg["h"], cfg["nfeat"]
    B = args.batch or min(cfg["batch"], 64)
    cap = bench.cap_of(cfg)
    base = [synth_frame(h, w, 7000 + i) for i in range(args.distinct)]           # identical on every rank (seeded)
    qmap = synth_descriptors(min(cfg["map_m"], 4096), 5)

    def make_frame(i):
        return np.roll(base[i % args.distinct], (3 * (i // args.distinct)) % w, axis=1)

    ctx = orb.Context(nf, bench.SCALE, bench.NLEVELS, w, h, B, device=local)

    def process(frames):
        kps, desc, cnt, best = ctx.extract_match_batch(frames, [qmap], cap)
        return [(kps[j, :cnt[j]], desc[j, :cnt[j]], best[0][j]) for j in range(len(frames))]

    process([make_frame(0)] * min(B, 4))                                          # warm-up (allocations, first launches)
    keep = sorted({int(x) for x in np.linspace(0, args.frames - 1, args.check)})
    if world > 1:
        dist.barrier()
    digests, kept, secs, op_secs = run_sharded_sequence(args.frames, make_frame, process, rank, world, args.mode, B, keep,
                                               device=torch.device("cuda", local))
    if rank == 0:
        bad = 0
        for i in keep:                                                            # the 1-GPU result of the same frame
            k, d, m = process([make_frame(i)])[0]
            bad += int(frame_digest(k, d, m) != digests[i])
            bad += int((k.tobytes(), d.tobytes(), m.tobytes()) != kept[i])
        print(json.dumps({"tool": "run_sequence", "config": args.config, "workload": bench.workload_string(cfg), "frames": args.frames, "n_gpus": world,
                          "sharding": args.mode, "frames_per_s_operator": args.frames / op_secs, "frames_per_s_whole_loop": args.frames / secs,
                          "seconds_in_operator_max_over_ranks": op_secs, "seconds_whole_loop_max_over_ranks": secs,
                          "frames_gathered_in_order": len(digests), "mean_keypoints": float(np.mean([g[0] for g in digests])),
                          "frames_compared_with_1gpu_result": keep, "mismatching_frames": bad,
                          "note": "host frames in pageable memory (numpy); operator = orbx_extract_match_batch calls (H2D, kernels, D2H inside); whole loop adds the synthetic "
                                  "frame production (numpy roll, the stand-in for decoding) and the digests; results gathered as per-frame digests + full records of the compared frames"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
