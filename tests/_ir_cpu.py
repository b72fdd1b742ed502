"""Test helper: executes the engine's op list (graph.py) on the CPU with torch so the
offset-write wiring can be compared with the conventionally written oracle."""
import numpy as np
import torch
import torch.nn.functional as F

from aerial_image_recognition_b200.graph import op_weights


def run_graph_cpu(g, weights, x_nchw, emulate_bf16=False):
    B = x_nchw.shape[0]
    bufs = {n: torch.zeros(B, b.c, b.h, b.w) for n, b in g.bufs.items()}
    rnd = (lambda t: t.to(torch.bfloat16).float()) if emulate_bf16 else (lambda t: t)
    bufs["input"][:, :3] = rnd(x_nchw)
    for op in g.ops:
        src = bufs[op.src.buf][:, op.src.c0:op.src.c0 + op.src.c]
        if op.kind in ("conv", "dwconv"):
            w, b = (torch.from_numpy(np.ascontiguousarray(a)).float() for a in op_weights(op, weights))
            groups = g.wshapes[op.weight][3]
            if op.src.buf == "input":
                src = src[:, :3]
            y = F.conv2d(src, w, b, stride=op.s, padding=op.k // 2, groups=groups)
            if op.act:
                y = y * torch.sigmoid(y)
            if op.res is not None:
                y = y + bufs[op.res.buf][:, op.res.c0:op.res.c0 + op.res.c]
            if not g.bufs[op.dst.buf].f32:
                y = rnd(y)
        elif op.kind == "maxpool":
            y = F.max_pool2d(src, op.k, op.s, op.k // 2 if op.s == 1 else 0)
        elif op.kind == "upsample2x":
            y = F.interpolate(src, scale_factor=2, mode="nearest")
        else:
            raise ValueError(op.kind)
        bufs[op.dst.buf][:, op.dst.c0:op.dst.c0 + op.dst.c] = y
    return bufs
