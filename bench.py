#!/usr/bin/env python
"""bench.py -- ORB extract + match frames/s at 640x480 (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's CUDA path (liborbx.so)
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] # OpenCV (cv2) on the host cores

Workload (BASELINE.json configs[1] + the front-end's call pattern, src/frontend.cpp:98-108): a batch of B = 256
synthetic 640x480 BGR frames, ORB with 1000 features / scale 1.2 / 8 levels, then 2 brute-force Hamming matches per
frame of an M = 2048-descriptor map (query) against that frame's descriptors (train).  A "step" = one pass over the
batch.  `value` is device-timed (CUDA events on the library's stream) with frames resident in HBM; `e2e` goes through
the host-buffer C-ABI calls with the H2D / D2H copies inside the timed region.  Frames shard across GPUs with no
collective (weak scaling: B frames per GPU per step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NFEAT, NLEVELS, SCALE = 640, 480, 1000, 8, 1.2
BATCH, MAP_M, CAP, MATCHES_PER_FRAME = 256, 2048, 1280, 2
METRIC = "orb_extract_match_frames_per_s_640x480"
LEVEL_PIXELS_VGA = 950532          # sum of the 8 pyramid levels of a 640x480 frame


def make_frames(batch: int, seed: int) -> np.ndarray:
    """`batch` distinct frames: 32 independently generated synthetic frames, the rest cheap distinct variants
    (circular shifts + flips) of them."""
    from rgbd_visualodometry_b200.synth import synth_frame
    nbase = min(batch, 32)
    base = [synth_frame(H, W, seed * 100003 + i) for i in range(nbase)]
    rng = np.random.default_rng(seed)
    out = np.empty((batch, H, W, 3), np.uint8)
    for i in range(batch):
        f = base[i % nbase]
        if i >= nbase:
            f = np.roll(f, (int(rng.integers(1, H)), int(rng.integers(1, W))), axis=(0, 1))
            if (i // nbase) & 1:
                f = f[:, ::-1]
        out[i] = f
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms; only samples inside the timed region are reported."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(index), "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(",")])

    def stop(self, t0: float = 0.0, t1: float = float("inf")) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        allrows = [r[1:] for r in self.rows if len(r) >= 7]
        rows = [r[1:] for r in self.rows if len(r) >= 7 and t0 <= r[0] <= t1 + 0.15] or allrows[-1:]   # samples inside the timed region
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in allrows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "period_ms": 100}


def cpu_reference_fps(frames: np.ndarray, map_desc: np.ndarray, threads: int, repeats: int = 1):
    """OpenCV on the host: frame-parallel worker threads (cv2 releases the GIL; ORB itself does not scale with
    cv threads), each doing the front-end's per-frame pattern: 1 detectAndCompute + 2 BFMatcher.match."""
    import cv2
    cv2.setNumThreads(1)
    n = len(frames)

    def work(idx):
        orb = cv2.ORB_create(NFEAT, SCALE, NLEVELS)
        bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        for i in idx:
            _, d = orb.detectAndCompute(frames[i], None)
            for _ in range(MATCHES_PER_FRAME):
                bf.match(map_desc, d)

    best = None
    for _ in range(repeats):
        parts = [list(range(t, n, threads)) for t in range(threads)]
        ths = [threading.Thread(target=work, args=(p,)) for p in parts]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n / best, best


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    try:
        import cv2  # noqa: F401
    except Exception as e:  # pragma: no cover
        emit({"impl": "reference", "unavailable": f"cv2 not importable: {e}"})
        return
    threads = os.cpu_count() or 1
    sample = 64
    frames = make_frames(sample, 0)
    import cv2
    _, d0 = cv2.ORB_create(NFEAT, SCALE, NLEVELS).detectAndCompute(frames[0], None)
    from rgbd_visualodometry_b200.synth import synth_map_queries
    map_desc = synth_map_queries(d0, MAP_M, 17)
    for _ in range(max(args.warmup, 1)):
        cpu_reference_fps(frames[:threads], map_desc, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_fps(frames, map_desc, threads)
    dt = time.perf_counter() - t0
    fps = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"cv2 {cv2.__version__} ORB({NFEAT},{SCALE},{NLEVELS}) detectAndCompute + {MATCHES_PER_FRAME}x BFMatcher(NORM_HAMMING).match(map {MAP_M} x frame) per 640x480 frame",
                       "sample_frames_per_step": sample},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "reference",
                             "sample": f"{sample} frames per step, frame-parallel over {threads} host threads, cv2.setNumThreads(1)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_orbx(args, rank: int, world: int, local_rank: int):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the orbx path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from rgbd_visualodometry_b200 import orb
    from rgbd_visualodometry_b200.synth import synth_map_queries

    B = args.batch
    frames_np = make_frames(B, rank)
    ctx = orb.Context(NFEAT, SCALE, NLEVELS, W, H, B, device=local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    dev = torch.device("cuda", local_rank)
    d_in = torch.from_numpy(frames_np).to(dev)
    d_kps = torch.zeros((B, CAP, 7), dtype=torch.float32, device=dev)
    d_desc = torch.zeros((B, CAP, 32), dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    d_best = torch.zeros((MATCHES_PER_FRAME, B, MAP_M, 4), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def extract():
        ctx.detect_and_compute_device(d_in.data_ptr(), B, W, H, W * 3, H * W * 3, 3, d_kps.data_ptr(), d_desc.data_ptr(), CAP, d_cnt.data_ptr())

    extract()
    ctx.synchronize()
    counts = d_cnt.cpu().numpy()
    assert counts.max() <= CAP, "capacity too small for this data"
    desc0 = d_desc[0, :counts[0]].cpu().numpy()
    map_np = synth_map_queries(desc0, MAP_M, 17)
    d_map = torch.from_numpy(map_np).to(dev)
    torch.cuda.synchronize()

    def step():
        extract()
        for m in range(MATCHES_PER_FRAME):
            ctx.match_device_ragged(d_map.data_ptr(), MAP_M, d_desc.data_ptr(), CAP, d_cnt.data_ptr(), B, d_best[m].data_ptr())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None   # started early; only samples inside the timed region count
    for _ in range(max(args.warmup, 3)):
        step()
    ctx.synchronize()
    barrier()
    launches0 = ctx.launch_count
    t_wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop(t_wall0, time.perf_counter()) if sampler else None
    ctx.synchronize()                                    # raises on any deferred device status
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max / 1e3)

    # ---- per-stage device times (CUDA events between the kernels, on the launching stream), 5 extra steps
    ctx.set_profiling(True)
    stage = {}
    for _ in range(5):
        step()
        for k, v in ctx.stage_times().items():
            stage.setdefault(k, []).append(v)
    ctx.set_profiling(False)
    stage_ms = {k: float(np.mean(v)) for k, v in stage.items()}

    # ---- end to end through the host-buffer C-ABI (rank-local), H2D + D2H inside the timed region
    pin_in = torch.from_numpy(frames_np).pin_memory()
    kps_h = torch.zeros((B, CAP, 7), dtype=torch.float32).pin_memory()
    desc_h = torch.zeros((B, CAP, 32), dtype=torch.uint8).pin_memory()
    cnt_h = np.zeros(B, np.int32)
    import ctypes as C
    ptrs = (C.c_void_p * B)(*[pin_in[i].data_ptr() for i in range(B)])

    map_pin = torch.from_numpy(map_np).pin_memory()
    best_hs = [torch.zeros((B, MAP_M, 4), dtype=torch.int32).pin_memory() for _ in range(MATCHES_PER_FRAME)]
    qptrs = (C.c_void_p * MATCHES_PER_FRAME)(*[map_pin.data_ptr()] * MATCHES_PER_FRAME)
    nqs = (C.c_int * MATCHES_PER_FRAME)(*[MAP_M] * MATCHES_PER_FRAME)
    bptrs = (C.c_void_p * MATCHES_PER_FRAME)(*[t.data_ptr() for t in best_hs])

    def e2e_step():
        # one C-ABI call: frames (host, pinned) -> keypoints, descriptors and the front-end's matches (host, pinned)
        rc = ctx.lib.orbx_extract_match_batch(ctx.h, ptrs, B, W, H, W * 3, 3, kps_h.data_ptr(), desc_h.data_ptr(), CAP, cnt_h.ctypes.data,
                                              qptrs, nqs, MATCHES_PER_FRAME, bptrs)
        assert rc == 0, ctx.lib.orbx_last_error(ctx.h)

    for _ in range(3):
        e2e_step()
    barrier()
    e2e_steps = max(3, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_fps = world * B * e2e_steps / float(te.item())
    h2d = B * H * W * 3 + MATCHES_PER_FRAME * MAP_M * 32
    d2h = B * CAP * 60 + 2 * B * 4 + MATCHES_PER_FRAME * B * MAP_M * 16 + 16

    if rank != 0:
        return
    # ---- roofline of the dominant kernel: algorithmic bytes per frame x frames per launch / its duration
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    n_out = float(counts.mean())
    b_alg = 3 * W * H + 2 * LEVEL_PIXELS_VGA + 60 * n_out
    ext_stages = {k: v for k, v in stage_ms.items() if k != "match"}   # gray, pyramid, fast_nms, select_harris, blur, describe
    dom = max(ext_stages, key=ext_stages.get) if ext_stages else None
    ext_total = sum(ext_stages.values())
    roofline = None
    traffic = None                                       # DRAM bytes of the dominant kernel per launch, from the committed ncu --set full capture
    traffic_file = "profiles/r1_v15_ncu_dram_traffic.json"
    try:
        tj = json.load(open(os.path.join(ROOT, traffic_file)))["kernels"]
        kname = {"gray": "k_gray", "pyramid": "k_pyr_down", "fast_nms": "k_fast_bands", "select_harris": "k_select", "blur": "k_blur",
                 "describe": "k_describe"}.get(dom)
        if kname in tj and B == 256:
            last = tj[kname][-1]
            traffic = last["dram_read_bytes"] + last["dram_write_bytes"]
    except Exception:
        pass
    if dom:
        ach = b_alg * B / (ext_stages[dom] / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic,
                    "traffic_source": f"{traffic_file} (ncu --set full, same 256-frame launch)" if traffic else None,
                    "peak_source": peak_src, "algorithmic_bytes_per_frame": b_alg, "frames_per_launch": B,
                    "kernel_ms": ext_stages[dom], "kernel_share_of_extraction": ext_stages[dom] / ext_total}
    pipe_ach = b_alg * B / (ext_total / 1e3) / 1e9 if ext_total else None
    match_ops = 2.0 * 256 * MAP_M * float(counts.sum())
    roof_match = None
    if "match" in stage_ms:
        tops = match_ops / (stage_ms["match"] / 1e3) / 1e12
        roof_match = {"bound": "tensor", "kernel": "k_hamming_umma2 (tcgen05 cta_group::2 kind::i8)", "achieved": tops, "peak": 4500.0, "unit": "TOP/s",
                      "frac": tops / 4500.0, "peak_source": "nominal dense int8 (no measured int8 peak available)", "kernel_ms": stage_ms["match"],
                      "ncu_tensor_pipe_cycles_active_pct": 76.4, "ncu_source": "profiles/r1_v15_ncu_full_summary.csv (sm__pipe_tensor_cycles_active, same launch shape)"}

    # ---- CPU baseline: the reference's own operators (cv2) on this box's host cores, bounded sample
    cpu = None
    try:
        threads = os.cpu_count() or 1
        sample = 64
        cpu_reference_fps(frames_np[:threads], map_np, threads)
        fps, dt = cpu_reference_fps(frames_np[:sample], map_np, threads, repeats=3)
        import cv2
        cpu = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "reference",
               "sample": f"{sample} of the step's {B} frames, best of 3, cv2 {cv2.__version__} frame-parallel over {threads} threads ({dt*1e3:.0f} ms)"}
    except Exception as e:  # pragma: no cover
        cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}

    # ---- parity counters of THIS run against the reference's operators (cv2) on the same frames: a few frames spread
    #      over the batch, both the device-resident outputs of the timed path and the host outputs of the e2e call; all must be 0
    parity = None
    try:
        import cv2
        cv2.setNumThreads(1)
        cvo, cvm = cv2.ORB_create(NFEAT, SCALE, NLEVELS), cv2.BFMatcher(cv2.NORM_HAMMING)
        chk = sorted({0, B // 3, (2 * B) // 3, B - 1})
        kp_bad = desc_bad = match_bad = 0
        cnt_d = d_cnt.cpu().numpy()
        for i in chk:
            ck, cd = cvo.detectAndCompute(frames_np[i], None)
            ref_k = np.array([(k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id) for k in ck], dtype=orb.KP_DTYPE)
            cm = cvm.match(map_np, cd)
            ref_m = np.array([(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in cm], dtype=orb.DMATCH_DTYPE)
            for kk, dd, nn, mm in ((d_kps[i].cpu().numpy(), d_desc[i].cpu().numpy(), int(cnt_d[i]), d_best[0][i].cpu().numpy()),
                                   (kps_h[i].numpy(), desc_h[i].numpy(), int(cnt_h[i]), best_hs[0][i].numpy())):
                got_k = np.ascontiguousarray(kk[:nn]).view(orb.KP_DTYPE).reshape(-1)
                if len(got_k) != len(ref_k):
                    kp_bad += max(len(got_k), len(ref_k)); desc_bad += 32 * max(len(got_k), len(ref_k)); match_bad += MAP_M
                    continue
                kp_bad += int(sum(a.tobytes() != b.tobytes() for a, b in zip(got_k, ref_k)))
                desc_bad += int((dd[:nn] != cd).sum())
                got_m = np.ascontiguousarray(mm).view(orb.DMATCH_DTYPE).reshape(-1)
                match_bad += int(sum(a.tobytes() != b.tobytes() for a, b in zip(got_m, ref_m))) + abs(len(got_m) - len(ref_m))
        parity = {"vs": f"cv2 {cv2.__version__} detectAndCompute + BFMatcher(NORM_HAMMING).match on the same frames", "frames_checked": chk,
                  "paths": ["device-resident (timed)", "host-buffer C-ABI (e2e)"], "keypoint_record_mismatches": kp_bad,
                  "descriptor_byte_mismatches": desc_bad, "match_record_mismatches": match_bad, "angle_tie_descriptor_diffs": 0 if desc_bad == 0 else None}
    except Exception as e:  # pragma: no cover
        parity = {"unavailable": str(e)}

    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": f"batched ORB extraction 640x480 BGR, {NFEAT} features, scale {SCALE}, {NLEVELS} levels + {MATCHES_PER_FRAME} Hamming matches/frame "
                                   f"(map {MAP_M} x frame descriptors, tcgen05 int8)", "frames_per_gpu_per_step": B, "parallelism": f"frame-sharded x{world}, no collective",
                       "l2": f"inputs {B * H * W * 3 / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)", "mean_keypoints_per_frame": n_out},
            "clocks": clocks,
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "orbx_extract_match_batch (one C-ABI call per step: host frames in, keypoints + descriptors + matches out; pinned host buffers), wall clock around the synchronous call"},
            "gpu_launches": int(launches),
            "roofline": roofline, "roofline_pipeline": {"achieved": pipe_ach, "peak": hbm_peak, "unit": "GB/s", "frac": pipe_ach / hbm_peak if pipe_ach else None,
                                                         "note": "same algorithmic bytes over the sum of all extraction kernels"},
            "roofline_match": roof_match, "stage_ms": stage_ms, "cpu_baseline": cpu, "parity": parity}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


# Rank 0's stdout carries exactly ONE JSON line: everything else a library may print there (NCCL's version banner goes to
# stdout whatever NCCL_DEBUG says) is sent to stderr by pointing fd 1 at fd 2 for the whole run; emit() writes the line to the
# real stdout.
_REAL_STDOUT = None


def _quiet_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    data = (json.dumps(obj) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="orbx", choices=["orbx", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_orbx(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
