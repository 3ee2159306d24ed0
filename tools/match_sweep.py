#!/usr/bin/env python
"""BASELINE config 3: frame<->map Hamming matching sweep, map size M = 1k..100k vs N = 2k frame descriptors, single set
and a 32-frame batched variant; device-resident, CUDA-event timed; reports int8 TOP/s against the nominal 4.5 POP/s.
Every result is checked against the CPU oracle on a sample of rows."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from rgbd_visualodometry_b200 import orb
from rgbd_visualodometry_b200.synth import synth_descriptors, synth_map_queries

N = 2000
ctx = orb.Context(1, 1.2, 1, 64, 64, 1)
st = torch.cuda.ExternalStream(ctx.stream)
print(f"{'M':>7s} {'sets':>4s} {'us/launch':>10s} {'TOP/s':>8s} {'% of 4.5 POP/s':>15s}  parity(sample)")
for nsets in (1, 32):
    trains = np.stack([synth_descriptors(N, 40 + s) for s in range(nsets)])
    dt = torch.from_numpy(trains).cuda()
    for M in (1000, 2000, 5000, 10000, 20000, 50000, 100000):
        q = synth_map_queries(trains[0], M, 41)
        dq = torch.from_numpy(q).cuda()
        best = torch.zeros((nsets, M, 4), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        for _ in range(3):
            ctx.match_device(dq.data_ptr(), M, dt.data_ptr(), N, nsets, best.data_ptr())
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            ctx.match_device(dq.data_ptr(), M, dt.data_ptr(), N, nsets, best.data_ptr())
        e1.record(st)
        ctx.synchronize()
        ms = e0.elapsed_time(e1) / reps
        ops = 2.0 * 256 * M * N * nsets
        got = best.cpu().numpy().view(orb.DMATCH_DTYPE).reshape(nsets, M)
        rows = np.random.default_rng(0).choice(M, min(M, 1500), replace=False)
        ok = all(np.array_equal(got[s][rows]["trainIdx"], O.match_hamming(q[rows], trains[s])["trainIdx"]) and
                 np.array_equal(got[s][rows]["distance"], O.match_hamming(q[rows], trains[s])["distance"]) for s in (0, nsets - 1))
        print(f"{M:7d} {nsets:4d} {ms*1e3:10.1f} {ops/ms/1e9:8.0f} {ops/ms/1e9/4500*100:14.1f}%  {'EXACT' if ok else 'MISMATCH'}")
