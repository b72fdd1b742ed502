"""Where the dedup stage of config C4 spends its time at world size N, measured on ONE GPU: the ranks run one after the
other (as in tests/test_gpu_parity.py::test_sharded_mosaic_equals_single_rank; the all-gather is a list), every phase of
``MosaicDetector.seam_split`` / ``finish`` bracketed by a synchronisation (``det.profile``), plus the un-instrumented
time of the same calls (CUDA events) and a checksum that must equal the single-rank result.

    python tools/c4_phases.py [--size 40000] [--world 8] [--ranks 0,3]
    B2D_VERBOSE=1 python tools/c4_phases.py ...        # also prints the round counts of the dedup / closure fixed points
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GT = (2335637.62, 0.1, 0.0, 6845688.78, 0.0, -0.1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=40000)
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--ranks", default="0,3", help="ranks whose finish() is timed")
    args = ap.parse_args()
    import torch
    from aerial_image_recognition_b200 import mosaic as M, synth
    from aerial_image_recognition_b200.engine import Engine

    H = W = args.size
    world = args.world
    eng = Engine("yolov8m", max_batch=64, device=0, seed=0)
    pool = torch.from_numpy(synth.mosaic_block_pool(77)).to(eng.device)
    det = M.MosaicDetector(eng, GT, conf=0.4, dedup_thr=1.0)
    covers = [M.shard_windows(H, W, r, world)[2] for r in range(world)]

    def ev():
        return torch.cuda.Event(enable_timing=True)

    locals_, recs, split_phases, split_ms, detect_ms, raw = [], [], [], [], [], []
    for r in range(world):
        wins, ids, cover = M.shard_windows(H, W, r, world)
        band = synth.mosaic_band_device(pool, H, W, cover[0], cover[1], 5)
        det.detect_windows(band, wins[:64], ids[:64], cover[0])                  # graph capture of the batch size
        a, b = ev(), ev()
        torch.cuda.synchronize()
        a.record()
        cols = det.detect_windows(band, wins, ids, cover[0])
        b.record()
        torch.cuda.synchronize()
        detect_ms.append(a.elapsed_time(b))
        raw.append(int(cols[0].numel()))
        del band
        # un-instrumented split, then the instrumented one
        det.profile = False
        a, b = ev(), ev()
        a.record()
        loc, rec = det.seam_split(*cols, r, covers, pack=False)
        b.record()
        torch.cuda.synchronize()
        split_ms.append(a.elapsed_time(b))
        det.profile, det.timings = True, {}
        det.seam_split(*cols, r, covers, pack=False)
        split_phases.append({k: round(v, 3) for k, v in det.timings.items()})
        det.profile = False
        locals_.append(loc)
        recs.append(rec)
        del cols
    origin = np.concatenate([np.full(len(rc), r, np.int64) for r, rc in enumerate(recs)])
    out = {"world": world, "size": H, "detect_ms_per_rank": [round(v, 2) for v in detect_ms], "raw_per_rank": raw,
           "seam_records_per_rank": [int(len(rc)) for rc in recs], "seam_split_ms_per_rank": [round(v, 2) for v in split_ms],
           "seam_split_phases_ms": split_phases, "finish": {}}
    total = 0
    ksum = 0
    for r in range(world):
        det.profile = False
        a, b = ev(), ev()
        t0 = time.perf_counter()
        a.record()
        res = det.finish(locals_[r], recs, origin, r)
        b.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        total += len(res)
        ksum += int(np.sum((res["window"].astype(np.int64) * 65536 + res["slot"].astype(np.int64)) % 1000003))
        if str(r) in args.ranks.split(","):
            det.profile, det.timings = True, {}
            det.finish(locals_[r], recs, origin, r)
            det.profile = False
            out["finish"][r] = {"ms_events": round(a.elapsed_time(b), 2), "ms_wall": round(wall, 2), "survivors": int(len(res)),
                                "phases_ms": {k: round(v, 3) for k, v in det.timings.items()}}
    out["detections_after_dedup"] = total
    out["keys_mod"] = ksum
    print(json.dumps(out))


if __name__ == "__main__":
    main()
