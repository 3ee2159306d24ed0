#!/usr/bin/env python
"""Stage-by-stage GPU-vs-oracle diagnostic (run on the B200 box).  Prints one line per stage and case with the
mismatch count, so one gpurun call localises a wrong kernel.  Not a test: always exits 0 unless it crashes."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import MATCH_CASES, ORB_CASES, ORB_CASES_LARGE  # noqa: E402
from oracle import oracle as O  # noqa: E402
from rgbd_visualodometry_b200.orb import Context  # noqa: E402


def diag_orb(name, img, n):
    h, w = img.shape[:2]
    ctx = Context(n, 1.2, 8, w, h, 2)
    try:
        t = time.time()
        kps, desc, cnt = ctx.detect_and_compute_batch([img, img], cap=max(2 * n, 64))
        dt = time.time() - t
        ko, do, dump = O.detect_and_compute(img, n, dump=True)
        ws, hs, sc, q = ctx.level_geometry(w, h)
        line = [f"{name}: gpu n={cnt.tolist()} oracle n={len(ko)} ({dt*1e3:.1f} ms)"]
        for l in range(8):
            if ws[l] <= 0 or hs[l] <= 0:
                continue
            lv = ctx.debug_level(0, l, int(ws[l]), int(hs[l]))
            bad = int((lv != dump["levels"][l]).sum())
            x, y, s = ctx.debug_fast(1, l)
            fo = dump["fast"][l]
            same_fast = len(x) == len(fo) and np.array_equal(x, fo["x"]) and np.array_equal(y, fo["y"]) and np.array_equal(s, fo["response"].astype(np.int32))
            line.append(f"  L{l} {ws[l]}x{hs[l]} pix_mismatch={bad} fast gpu={len(x)} oracle={len(fo)} {'OK' if same_fast else 'FAST-MISMATCH'}")
            if not same_fast and len(fo):
                so = set(zip(fo["x"].tolist(), fo["y"].tolist(), fo["response"].astype(int).tolist()))
                sg = set(zip(x.tolist(), y.tolist(), s.tolist()))
                line.append(f"     only_gpu={len(sg - so)} only_oracle={len(so - sg)} e.g. {sorted(sg - so)[:3]} / {sorted(so - sg)[:3]}")
        print("\n".join(line))
        for b in range(2):
            k = kps[b, :cnt[b]]
            d = desc[b, :cnt[b]]
            if len(k) != len(ko):
                print(f"  frame{b}: COUNT MISMATCH")
                continue
            res = {f: int((k[f] != ko[f]).sum()) for f in k.dtype.names}
            drows = int((d != do).any(axis=1).sum())
            pos_set = set(zip(k["x"].tolist(), k["y"].tolist(), k["octave"].tolist())) == set(zip(ko["x"].tolist(), ko["y"].tolist(), ko["octave"].tolist()))
            print(f"  frame{b}: field mismatches {res} desc_rows_diff={drows} same_set={pos_set} -> {'EXACT' if not any(res.values()) and drows == 0 else 'DIFF'}")
    finally:
        ctx.close()


def diag_match(name, q, t):
    ctx = Context(1, 1.2, 1, 64, 64, 1)
    try:
        mo = O.match_hamming(q, t)
        t0 = time.time()
        mg = ctx.match(q, t)
        dt = time.time() - t0
        ok = mg.tobytes() == mo.tobytes()
        msg = f"match {name}: {q.shape}x{t.shape} {'EXACT' if ok else 'DIFF'} ({dt*1e3:.1f} ms)"
        if not ok and len(mg) == len(mo):
            msg += f" idx_diff={(mg['trainIdx'] != mo['trainIdx']).sum()} dist_diff={(mg['distance'] != mo['distance']).sum()} first gpu={mg[:3].tolist()} oracle={mo[:3].tolist()}"
        print(msg)
        if len(t) >= 1:
            ko = O.match_hamming_knn2(q, t)
            kg = ctx.knn_match2(q, t)
            print(f"knn2  {name}: {'EXACT' if kg.tobytes() == ko.tobytes() else 'DIFF'}")
    finally:
        ctx.close()


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "orb"):
        for name, (mk, n) in list(ORB_CASES.items()) + list(ORB_CASES_LARGE.items()):
            try:
                diag_orb(name, mk(), n)
            except Exception:
                print(f"{name}: EXCEPTION"); traceback.print_exc()
            sys.stdout.flush()
    if which in ("all", "match"):
        for name, (mq, mt) in MATCH_CASES.items():
            try:
                diag_match(name, mq(), mt())
            except Exception:
                print(f"match {name}: EXCEPTION"); traceback.print_exc()
            sys.stdout.flush()
        from rgbd_visualodometry_b200.synth import synth_descriptors, synth_map_queries
        t = synth_descriptors(2000, 40)
        for m in (1000, 10000, 100000):
            try:
                diag_match(f"sweep_{m}", synth_map_queries(t, m, 41), t)
            except Exception:
                print(f"match sweep_{m}: EXCEPTION"); traceback.print_exc()
            sys.stdout.flush()


if __name__ == "__main__":
    main()
