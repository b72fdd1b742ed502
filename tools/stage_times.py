"""Per-stage device times of one inference step (preprocess / forward / postprocess / georef) at batch 64."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from aerial_image_recognition_b200 import synth
from aerial_image_recognition_b200.engine import Engine
B = 64
eng = Engine("yolov8m", max_batch=B)
t = torch.from_numpy(synth.make_tiles(8, 640, 5)).cuda().repeat(8, 1, 1, 1).contiguous()
params = torch.zeros((B, 16), dtype=torch.float64, device="cuda"); params[:, :6] = torch.tensor([21.0, 21.00094, 52.2, 52.200575, 864.0, 640.0], dtype=torch.float64)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
acc = np.zeros(4)
for it in range(8):
    ev[0].record(); eng.preprocess(t, "identity")
    ev[1].record(); eng.forward(B)
    ev[2].record(); dets, counts = eng.postprocess(B, 0.25, False, 0.7, 0, 300, None)
    ev[3].record(); g = eng.georef(dets, counts, params, "bounds")
    ev[4].record(); torch.cuda.synchronize()
    if it >= 3: acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(4)])
acc /= 5
print("preprocess %.3f ms | forward %.3f ms | postprocess %.3f ms | georef %.3f ms | total %.3f ms" % (*acc, acc.sum()))
print("detections", int(counts.sum()))
