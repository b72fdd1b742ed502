"""Run one planned op repeatedly (for ncu): python tools/one_op.py --op 9 --reps 3 [--batch 64]"""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aerial_image_recognition_b200 import synth
from aerial_image_recognition_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--op", type=int, nargs="+", default=[9])
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--arch", default="yolov8m")
a = ap.parse_args()
eng = Engine(a.arch, max_batch=a.batch)
t = torch.from_numpy(synth.make_tiles(4, eng.imgsz, 5)).cuda().repeat(a.batch // 4, 1, 1, 1).contiguous()
eng.preprocess(t, "identity")
eng.forward(a.batch)
torch.cuda.synchronize()
for op in a.op:
    print(eng.describe_op(op))
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.reps):
        eng.run_op(op, a.batch)
    e.record(); torch.cuda.synchronize()
    print(f"op {op}: {s.elapsed_time(e)/a.reps*1e3:.1f} us")
