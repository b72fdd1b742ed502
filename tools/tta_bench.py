#!/usr/bin/env python
"""Times the test-time-augmentation view kernels (csrc/tta.cu) and the five-view detection step on one B200.

    python tools/tta_bench.py [--batch 64] [--iters 20] > gpurun_out/tta_bench.json

Per kernel family: CUDA-event time per call on a batch of 640x640 tiles resident in HBM (inputs 79 MB per batch of 64,
cycled over 4 batches so the 126 MB L2 does not hold them), algorithmic bytes (reads + writes of the 3-byte pixels)
and the fraction of the measured HBM peak.  Then the whole five-view step of GPUHandler.process_batch_tta's device part
(views -> preprocess -> network -> scaled strict filter -> float32 georef), in tiles/s.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aerial_image_recognition_b200 import synth, tta as T   # noqa: E402
from aerial_image_recognition_b200.engine import GEO_PARAMS, Engine   # noqa: E402


def timed(fn, iters, warm=3):
    for k in range(warm):
        fn(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(iters):
        fn(k)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def cpu_views(ntiles):
    """The reference's own view code (gpu_handler.py:94-140: cv2 + Pillow + NumPy on the host) timed on `ntiles` tiles."""
    import time
    import cv2
    from PIL import Image, ImageEnhance
    tiles = synth.make_tiles(ntiles, 640, 300)
    cv2.setNumThreads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    for a in tiles:
        img = Image.fromarray(a)
        for clip, grid in ((3.0, 8), (4.0, 4)):
            lab = cv2.cvtColor(a, cv2.COLOR_RGB2LAB)
            l, aa, bb = cv2.split(lab)
            le = cv2.createCLAHE(clipLimit=clip, tileGridSize=(grid, grid)).apply(l)
            cv2.cvtColor(cv2.merge([le, aa, bb]), cv2.COLOR_LAB2RGB)
        np.array(ImageEnhance.Brightness(img).enhance(2.0))
        (np.power(a / 255.0, 1.0 / 2.0) * 255.0).astype(np.uint8)
    dt = time.perf_counter() - t0
    return {"value": round(ntiles / dt, 1), "unit": "tiles/s (four derived views per tile, no network)", "cores": os.cpu_count(),
            "kind": "reference", "sample": f"{ntiles} synthetic 640x640 tiles through the reference's cv2 / Pillow / NumPy lines"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--cpu-tiles", type=int, default=128)
    args = ap.parse_args()
    n = args.batch
    peak = 6532.2
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    eng = Engine("yolov8m", max_batch=n, seed=0)
    pool = [torch.from_numpy(synth.make_tiles(n, 640, 100 + k)).cuda() for k in range(4)]
    img_bytes = 640 * 640 * 3
    bright, gamma = T.brightness_lut(2.0), T.gamma_lut(2.0)
    out = {"batch": n, "hbm_peak_gbps": peak, "kernels": {}}

    def report(name, ms, passes):
        gb = passes * n * img_bytes / 1e9
        out["kernels"][name] = {"ms_per_batch": round(ms, 4), "algorithmic_GB": round(gb, 4), "GBps": round(gb / ms * 1e3, 1),
                                "frac_of_hbm_peak": round(gb / ms * 1e3 / peak, 3)}
    report("clahe(3.0, 8x8): lab histogram pass + lab/interp/rgb pass", timed(lambda k: eng.tta_clahe(pool[k % 4], 3.0, 8), args.iters), 3)
    report("clahe(4.0, 4x4)", timed(lambda k: eng.tta_clahe(pool[k % 4], 4.0, 4), args.iters), 3)
    report("byte curve (brightness 2.0)", timed(lambda k: eng.tta_lut(pool[k % 4], bright), args.iters), 2)
    report("contrast 1.3: grey-mean pass + curve pass", timed(lambda k: eng.tta_contrast(pool[k % 4], 1.3), args.iters), 3)
    report("rgb2lab", timed(lambda k: eng.colour_convert(pool[k % 4], "rgb2lab"), args.iters), 2)
    report("lab2rgb", timed(lambda k: eng.colour_convert(pool[k % 4], "lab2rgb"), args.iters), 2)

    views = T.LIGHTING_VIEWS + T.OCCLUSION_VIEWS
    params = torch.zeros((n, GEO_PARAMS), dtype=torch.float64, device="cuda")
    params[:, :4] = torch.tensor([21.0, 52.0, 21.0006, 52.0004], dtype=torch.float64)

    def step(k):
        for i, v in enumerate(eng.tta_views(pool[k % 4], views)):
            dets, counts = eng.infer(v, "identity", True, 0.3, False, conf_scale=T.confidence_adjustment(i))
            eng.georef(dets, counts, params, "tensor_f32")
    ms = timed(step, max(5, args.iters // 2))
    plain = timed(lambda k: eng.georef(*eng.infer(pool[k % 4], "identity", False, 0.3, True), params, "gpuhandler"), args.iters)
    vms = timed(lambda k: eng.tta_views(pool[k % 4], views), args.iters)
    out["views_only"] = {"ms_per_batch": round(vms, 4), "tiles_per_s": round(n / vms * 1e3, 1), "note": "the four derived views of every tile"}
    out["five_view_step"] = {"ms_per_batch": round(ms, 3), "tiles_per_s": round(n / ms * 1e3, 1), "views": len(views),
                             "network_passes_per_s": round(len(views) * n / ms * 1e3, 1)}
    out["single_view_step"] = {"ms_per_batch": round(plain, 3), "tiles_per_s": round(n / plain * 1e3, 1)}
    out["cpu_baseline"] = cpu_views(args.cpu_tiles)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
