"""Seeded synthetic deploy-form weights for the detector graphs.

Every model blob the reference loads is absent (``.MISSING_LARGE_BLOBS:2-5``;
``_script/config.py:25``, ``simple_detector.py:710``), and there is no network,
so throughput and parity are measured with random-init weights of the right
architecture.  The generator is deterministic in ``(arch, nc, seed)``.

* conv weights are He-style for SiLU (pre-activation variance ~1), then rounded
  to bf16-representable fp32 so the bf16 engine and the fp32 oracle see *the
  same* weights (what is compared is the arithmetic, not a quantiser);
* BatchNorm is already folded: each conv has ``.weight`` and ``.bias``;
* the last cls / objectness bias is shifted so that a few percent of anchors
  clear the reference's 0.3 threshold (``simple_detector.py:30``), and the DFL
  logits are tilted towards small bins so boxes are car-sized rather than
  tile-sized.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from .graph import Graph, op_weights

def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (NaN-free inputs)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(x.shape)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return u.astype(np.uint16).reshape(x.shape)


def _quantised_scale(s: float) -> float:
    """Snap a calibration scale to 2**(j/8) so last-bit differences in the CPU
    conv used for calibration cannot change the generated weights."""
    return float(2.0 ** (np.round(np.log2(max(s, 1e-12)) * 8.0) / 8.0))


def make_synthetic_weights(graph: Graph, seed: int = 0, calibrate: bool = True) -> Dict[str, np.ndarray]:
    """Deterministic deploy-form weights for ``graph``.

    With ``calibrate`` (default) each conv is rescaled layer by layer, in graph
    order, so that its pre-activation has unit standard deviation on one seeded
    synthetic 320x320 tile (LSUV-style).  A fixed analytic gain does not work for
    SiLU: the per-layer gain depends on the signal scale, so depth-60 stacks
    either collapse onto the bias noise or overflow.
    """
    rng = np.random.default_rng([seed, 0xB200])
    w: Dict[str, np.ndarray] = {}
    final = set()
    if graph.head["kind"] == "v8_dfl":
        for i in range(3):
            final.add(f"model.22.cv2.{i}.2")
            final.add(f"model.22.cv3.{i}.2")
    else:
        for i in range(3):
            final.add(f"model.105.m.{i}")
    for name, (cout, cing, k, groups) in graph.wshapes.items():
        fan_in = cing * k * k
        std = (2.0 / fan_in) ** 0.5
        wt = rng.standard_normal((cout, cing, k, k), dtype=np.float32) * np.float32(std)
        b = rng.standard_normal(cout, dtype=np.float32) * np.float32(0.1)
        if name in final:
            if "cv2" in name and graph.head["kind"] == "v8_dfl":
                # DFL logits: tilt towards small bins -> boxes of a few cells
                b = b + np.tile(-0.6 * np.arange(16, dtype=np.float32), 4)
            elif graph.head["kind"] == "v8_dfl":
                b = b - np.float32(3.0)
            else:
                no = graph.nc + 5
                bb = b.reshape(3, no)
                bb[:, 4] -= np.float32(3.0)
                b = bb.reshape(-1)
        w[name + ".weight"] = round_to_bf16(wt)
        w[name + ".bias"] = b.astype(np.float32)
    if calibrate:
        _lsuv(graph, w, seed)
    return w


def _lsuv(graph: Graph, w: Dict[str, np.ndarray], seed: int) -> None:
    """Layer-sequential unit-variance rescale on a small calibration tile (CPU torch;
    this is weight *generation*, not the inference path)."""
    import torch
    import torch.nn.functional as F
    from .graph import build
    from .synth import make_tiles

    hw = 320
    g = build(graph.arch, graph.nc, hw)
    x = torch.from_numpy(make_tiles(1, hw, seed=seed + 7919)[0].astype(np.float32) / 255.0).permute(2, 0, 1)[None]
    bufs = {n: torch.zeros(1, b.c, b.h, b.w) for n, b in g.bufs.items()}
    bufs["input"][:, :3] = x
    with torch.no_grad():
        for op in g.ops:
            src = bufs[op.src.buf][:, op.src.c0:op.src.c0 + op.src.c]
            if op.kind in ("conv", "dwconv"):
                if op.src.buf == "input":
                    src = src[:, :3]
                wt, b = (torch.from_numpy(np.ascontiguousarray(a)) for a in op_weights(op, w))    # buffer channel order
                groups = g.wshapes[op.weight][3]
                y = F.conv2d(src, wt, None, stride=op.s, padding=op.k // 2, groups=groups)
                sc = _quantised_scale(1.0 / float(y.std()))
                wt = torch.from_numpy(round_to_bf16((wt * sc).numpy()))
                w[op.weight + ".weight"] = round_to_bf16((torch.from_numpy(w[op.weight + ".weight"]) * sc).numpy())   # model order
                y = F.conv2d(src, wt, b, stride=op.s, padding=op.k // 2, groups=groups)
                if op.act:
                    y = y * torch.sigmoid(y)
                if op.res is not None:
                    y = y + bufs[op.res.buf][:, op.res.c0:op.res.c0 + op.res.c]
            elif op.kind == "maxpool":
                y = F.max_pool2d(src, op.k, op.s, op.k // 2 if op.s == 1 else 0)
            else:
                y = F.interpolate(src, scale_factor=2, mode="nearest")
            bufs[op.dst.buf][:, op.dst.c0:op.dst.c0 + op.dst.c] = y


def weights_fingerprint(w: Dict[str, np.ndarray]) -> str:
    import hashlib
    h = hashlib.sha256()
    for k in sorted(w):
        h.update(k.encode())
        h.update(np.ascontiguousarray(w[k]).tobytes())
    return h.hexdigest()[:16]
