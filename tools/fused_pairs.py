"""Time every fused depthwise + pointwise pair against its two single-op kernels: python tools/fused_pairs.py [--batch 64] [--reps 20]"""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aerial_image_recognition_b200 import synth
from aerial_image_recognition_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--arch", default="yolov8m")
ap.add_argument("--only", type=int, default=-1, help="time only the pair starting at this op, fused kernel only (for ncu)")
a = ap.parse_args()
eng = Engine(a.arch, max_batch=a.batch)
t = torch.from_numpy(synth.make_tiles(4, 640, 5)).cuda().repeat(a.batch // 4, 1, 1, 1).contiguous()
eng.preprocess(t, "identity")
eng.forward(a.batch)
torch.cuda.synchronize()

def timed(fn):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.reps):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / a.reps * 1e3

for i in range(len(eng.graph.ops) - 1):
    if not eng.fused_with_next(i) or (a.only >= 0 and i != a.only):
        continue
    if a.only >= 0:
        print(f"ops {i}+{i + 1}: fused {timed(lambda: eng.run_op_fused(i, a.batch)):7.1f} us")
        continue
    f = timed(lambda: eng.run_op_fused(i, a.batch))
    d = timed(lambda: eng.run_op(i, a.batch))
    p = timed(lambda: eng.run_op(i + 1, a.batch))
    print(f"ops {i}+{i + 1}: fused {f:7.1f} us | depthwise {d:6.1f} + pointwise {p:6.1f} = {d + p:7.1f} us | {eng.describe_op(i).split('|')[0].strip()}")
