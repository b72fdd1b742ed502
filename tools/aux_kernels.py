#!/usr/bin/env python
"""One pass of the stages either side of the network at batch 64 (for `ncu -k regex:...` captures of the pre-/post-processing
and view kernels):  python tools/aux_kernels.py [--reps 2]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aerial_image_recognition_b200 import synth   # noqa: E402
from aerial_image_recognition_b200.engine import GEO_PARAMS, Engine   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--batch", type=int, default=64)
a = ap.parse_args()
eng = Engine("yolov8m", max_batch=a.batch, seed=0)
tiles = torch.from_numpy(synth.make_tiles(a.batch, 640, 7)).cuda()
params = torch.zeros((a.batch, GEO_PARAMS), dtype=torch.float64, device="cuda")
params[:, :6] = torch.tensor([21.0, 21.0006, 52.0, 52.0004, 864, 640], dtype=torch.float64)
for _ in range(a.reps):
    eng.preprocess(tiles, "identity")
    eng.forward(a.batch)
    dets, counts = eng.postprocess(a.batch, 0.25, False, 0.7, 0, 300)
    eng.georef(dets, counts, params, "bounds")
    eng.tta_clahe(tiles, 3.0, 8)
torch.cuda.synchronize()
print("done", int(counts.sum()))
