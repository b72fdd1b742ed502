"""Step time and SM clock over a long back-to-back run of the C2 step (is `value` a steady-state number?).
    python tools/steady_state.py [--chunks 40] [--steps 10] [--idle 0]"""
import argparse, os, subprocess, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from aerial_image_recognition_b200 import synth
from aerial_image_recognition_b200.engine import Engine
ap = argparse.ArgumentParser(); ap.add_argument("--chunks", type=int, default=40); ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--idle", type=float, default=0.0); ap.add_argument("--precision", default="bf16"); ap.add_argument("--seed", type=int, default=0)
a = ap.parse_args()
B = 64
eng = Engine("yolov8m", max_batch=B, precision=a.precision, seed=a.seed)
base = synth.make_tiles(16, 640, seed=1000)
pool = [torch.from_numpy(base[(np.arange(B) * 7 + b * 3) % 16]).cuda() for b in range(4)]
params = torch.zeros((B, 16), dtype=torch.float64, device="cuda"); params[:, :6] = torch.tensor([21.0, 21.00094, 52.2, 52.200575, 864.0, 640.0], dtype=torch.float64)
rows = []
def sample():
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        rows.append((time.perf_counter(), line.strip()))
threading.Thread(target=sample, daemon=True).start()
def step(i):
    dets, counts = eng.infer(pool[i % 4], "identity", False, 0.25, False, 0.7, 0, 300)
    return eng.georef(dets, counts, params, "bounds")
for i in range(3): step(i)
torch.cuda.synchronize()
if a.idle: time.sleep(a.idle)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.chunks + 1)]
t0 = time.perf_counter()
ev[0].record()
for c in range(a.chunks):
    for i in range(a.steps): step(c * a.steps + i)
    ev[c + 1].record()
torch.cuda.synchronize()
t1 = time.perf_counter()
ms = [ev[c].elapsed_time(ev[c + 1]) / a.steps for c in range(a.chunks)]
print("ms/step per chunk:", " ".join(f"{m:.2f}" for m in ms))
print("clock/power/temp samples during run:", " | ".join(r[1] for r in rows if t0 <= r[0] <= t1))
