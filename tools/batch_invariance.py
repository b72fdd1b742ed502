"""Are a tile's detections independent of the batch it runs in?  (n = 64 vs 63 / 57 / 33 / 24 / 7 / 1, stale neighbours)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from aerial_image_recognition_b200 import synth
from aerial_image_recognition_b200.engine import Engine, dets_to_numpy
eng = Engine("yolov8m", max_batch=64)
t = torch.from_numpy(synth.make_tiles(64, 640, 77)).cuda()
other = torch.from_numpy(synth.make_tiles(64, 640, 78)).cuda()
def run(x):
    d, c = eng.infer(x, "identity", False, 0.25, False, 0.7, 0, 300)
    return dets_to_numpy(d, c)
ref = run(t)
for n in (63, 57, 33, 24, 7, 1):
    for rep in range(3):                      # eager, captured, replayed
        run(other)                            # stale data of another batch in every buffer
        got = run(t[:n])
        bad = [i for i in range(n) if not (len(got[i]) == len(ref[i]) and all(np.array_equal(got[i][f], ref[i][f]) for f in ("cx", "cy", "w", "h", "conf", "cls", "anchor")))]
        print(f"n={n} call {rep}: {len(bad)} tiles differ", bad[:8])
        if bad:
            i = bad[0]
            print("   e.g. tile", i, "counts", len(got[i]), len(ref[i]), "max |d conf|", np.abs(got[i]["conf"][:min(len(got[i]), len(ref[i]))] - ref[i]["conf"][:min(len(got[i]), len(ref[i]))]).max())
