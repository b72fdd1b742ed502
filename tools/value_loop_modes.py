"""Is the back-to-back step time bimodal?  Repeats of 50-step loops in one process, with and without periodic host syncs."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from aerial_image_recognition_b200 import synth
from aerial_image_recognition_b200.engine import Engine
B = 64
eng = Engine("yolov8m", max_batch=B)
base = synth.make_tiles(16, 640, seed=1000)
pool = [torch.from_numpy(base[(np.arange(B) * 7 + b * 3) % 16]).cuda() for b in range(4)]
params = torch.zeros((B, 16), dtype=torch.float64, device="cuda"); params[:, :6] = torch.tensor([21.0, 21.00094, 52.2, 52.200575, 864.0, 640.0], dtype=torch.float64)
def step(i):
    dets, counts = eng.infer(pool[i % 4], "identity", False, 0.25, False, 0.7, 0, 300)
    return eng.georef(dets, counts, params, "bounds")
def loop(k, sync_every=0, fwd_only=False):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(k):
        if fwd_only:
            eng.forward(B)
        else:
            step(i)
        if sync_every and (i + 1) % sync_every == 0:
            torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / k
for i in range(200): step(i)
torch.cuda.synchronize()
for rep in range(4):
    print("plain %.2f | sync/10 %.2f | sync/1 %.2f | forward only %.2f | plain %.2f" % (loop(50), loop(50, 10), loop(50, 1), loop(50, 0, True), loop(50)), flush=True)
