#!/usr/bin/env python
"""Throughput of BASELINE.json configs 2, 4 and 5 on ONE GPU (device-resident frames), with the SURVEY 8(d) roofline
figures: frames/s, algorithmic GB/s against the measured HBM peak, matcher TOP/s, and a parity spot check of one frame
against the oracle.  (bench.py is the contract line for config 2; this tool is the per-config report.)

  python tools/config_bench.py [c2] [c4] [c5]
"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgbd_visualodometry_b200 import orb
from rgbd_visualodometry_b200.synth import synth_frame, synth_map_queries

CONFIGS = {   # name: (W, H, nfeatures, batch, distinct frames, map size, level pixels P)
    "c2": (640, 480, 1000, 256, 32, 2048, 950532),
    "c4": (1920, 1080, 2000, 64, 8, 10000, 6419321),
    "c5": (3840, 2160, 5000, 64, 4, 10000, 25677702),
}
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)


def run(name):
    W, H, NF, B, ND, M, P = CONFIGS[name]
    base = [synth_frame(H, W, 777 + i) for i in range(ND)]
    frames = np.stack([np.roll(base[i % ND], (7 * (i // ND), 13 * (i // ND)), axis=(0, 1)) for i in range(B)])
    cap = int(NF * 1.3) + 64
    ctx = orb.Context(NF, 1.2, 8, W, H, B)
    d_in = torch.from_numpy(frames).cuda()
    d_k = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda")
    d_d = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(B, dtype=torch.int32, device="cuda")
    st = torch.cuda.ExternalStream(ctx.stream)

    def extract():
        ctx.detect_and_compute_device(d_in.data_ptr(), B, W, H, W * 3, H * W * 3, 3, d_k.data_ptr(), d_d.data_ptr(), cap, d_n.data_ptr())

    extract(); ctx.synchronize()
    cnt = d_n.cpu().numpy()
    assert cnt.max() <= cap
    # parity spot check: frame 0 against the oracle (bit-exact)
    from oracle import oracle as O
    ko, do = O.detect_and_compute(frames[0], NF)
    k0 = d_k[0, :cnt[0]].cpu().numpy().tobytes(); d0 = d_d[0, :cnt[0]].cpu().numpy()
    parity = "EXACT" if (k0 == ko.tobytes() and np.array_equal(d0, do)) else "MISMATCH"
    d_map = torch.from_numpy(synth_map_queries(d0, M, 5)).cuda()
    d_best = torch.zeros((B, M, 4), dtype=torch.int32, device="cuda")

    def match():
        ctx.match_device_ragged(d_map.data_ptr(), M, d_d.data_ptr(), cap, d_n.data_ptr(), B, d_best.data_ptr())

    res = {}
    for what, fn, reps in (("extract", extract, 20), ("match", match, 20)):
        for _ in range(3): fn()
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps): fn()
        e1.record(st); ctx.synchronize()
        res[what] = e0.elapsed_time(e1) / reps
    n_out = float(cnt.mean())
    b_alg = 3 * W * H + 2 * P + 60 * n_out
    gbs = b_alg * B / (res["extract"] / 1e3) / 1e9
    tops = 2.0 * 256 * M * float(cnt.sum()) / (res["match"] / 1e3) / 1e12
    print(f"{name}: {W}x{H} nfeatures={NF} batch={B}  extract {res['extract']:.3f} ms = {B / res['extract'] * 1e3:,.0f} frames/s  "
          f"algorithmic {gbs:,.0f} GB/s = {gbs / peak * 100:.1f}% of measured {peak:.0f} GB/s | match vs {M}-row map {res['match']:.3f} ms = "
          f"{tops:,.0f} TOP/s = {tops / 4500 * 100:.1f}% of nominal int8 | extract+match {B / (res['extract'] + res['match']) * 1e3:,.0f} frames/s | "
          f"mean keypoints {n_out:.0f} | parity(frame 0 vs oracle) {parity}", flush=True)
    ctx.close()


for name in (sys.argv[1:] or ["c2", "c4", "c5"]):
    run(name)
