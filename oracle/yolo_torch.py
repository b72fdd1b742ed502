"""ORACLE (test infrastructure, not product code) -- PyTorch fp32 CPU restatement
of the detector networks the reference runs through onnxruntime.

PARITY UNPINNED: the reference holds no golden vectors, no tests and no model
blobs for this path (``.MISSING_LARGE_BLOBS:2-5``), and
``onnxruntime`` / ``ultralytics`` are not installable here.  This file restates
the *published* module semantics of Ultralytics 8.3.4 YOLOv8 (the architecture
logged at ``x_arch/01_train_tokyo.ipynb:1 (cell 15 output)``) and of the
canonical YOLOv7 deploy graph, and is anchored on the reference's call sites
``simple_detector.py:474`` / ``_script/gpu_handler.py:165`` (input
``float32[1,3,640,640]`` in [0,1], RGB, output rows / channels as documented in
SURVEY.md section 8a rows a4-a5).  It is written the conventional way (NCHW, explicit
``torch.cat`` / ``chunk``) on purpose, so that it is independent of the engine's
offset-write graph in ``aerial_image_recognition_b200/graph.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import this module.

``emulate_bf16=True`` rounds every stored activation to bf16 exactly where the
engine stores bf16 (after bias+SiLU(+residual) of each conv; not the network
input, which the engine keeps as the exact 8-bit pixel value); accumulation stays fp32.  That mode checks the kernels' arithmetic;
``emulate_bf16=False`` is the fp32 stand-in for the onnxruntime CPU path.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F


def _bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


class _Net:
    def __init__(self, weights: Dict[str, np.ndarray], emulate_bf16: bool):
        self.w = {k: torch.from_numpy(np.ascontiguousarray(v)).float() for k, v in weights.items()}
        self.emu = emulate_bf16

    def rnd(self, x):
        # emulate_bf16: True / "bf16" rounds stored activations to bf16, "fp16" to fp16 (the engine's B2D_PREC_FP16 mode)
        if self.emu == "fp16":
            return x.to(torch.float16).to(torch.float32)
        return _bf16(x) if self.emu else x

    def conv(self, name, x, k=1, s=1, act=True, res=None, groups=1, out_f32=False):
        w, b = self.w[name + ".weight"], self.w[name + ".bias"]
        assert w.shape[-1] == k, (name, w.shape, k)
        y = F.conv2d(x, w, b, stride=s, padding=k // 2, groups=groups)
        if act:
            y = y * torch.sigmoid(y)
        if res is not None:
            y = y + res
        return y if out_f32 else self.rnd(y)


class YoloV8mOracle(_Net):
    """Ultralytics 8.3.4 YOLOv8m, nc classes, deploy (BN-folded) form."""

    def c2f(self, name, x, n, shortcut):
        y = list(self.conv(f"{name}.cv1", x).chunk(2, 1))
        for i in range(n):
            t = self.conv(f"{name}.m.{i}.cv1", y[-1], 3)
            y.append(self.conv(f"{name}.m.{i}.cv2", t, 3, res=y[-1] if shortcut else None))
        return self.conv(f"{name}.cv2", torch.cat(y, 1))

    def sppf(self, name, x):
        y = [self.conv(f"{name}.cv1", x)]
        for _ in range(3):
            y.append(F.max_pool2d(y[-1], 5, 1, 2))
        return self.conv(f"{name}.cv2", torch.cat(y, 1))

    @torch.no_grad()
    def raw_head(self, x: torch.Tensor) -> List[torch.Tensor]:
        """x: [B,3,H,W] fp32 in [0,1].  Returns per level [B, 64+nc, h, w] fp32 raw."""
        x0 = self.conv("model.0", x, 3, 2)
        x1 = self.conv("model.1", x0, 3, 2)
        x2 = self.c2f("model.2", x1, 2, True)
        x3 = self.conv("model.3", x2, 3, 2)
        x4 = self.c2f("model.4", x3, 4, True)
        x5 = self.conv("model.5", x4, 3, 2)
        x6 = self.c2f("model.6", x5, 4, True)
        x7 = self.conv("model.7", x6, 3, 2)
        x8 = self.c2f("model.8", x7, 2, True)
        x9 = self.sppf("model.9", x8)
        x12 = self.c2f("model.12", torch.cat([F.interpolate(x9, scale_factor=2, mode="nearest"), x6], 1), 2, False)
        x15 = self.c2f("model.15", torch.cat([F.interpolate(x12, scale_factor=2, mode="nearest"), x4], 1), 2, False)
        x16 = self.conv("model.16", x15, 3, 2)
        x18 = self.c2f("model.18", torch.cat([x16, x12], 1), 2, False)
        x19 = self.conv("model.19", x18, 3, 2)
        x21 = self.c2f("model.21", torch.cat([x19, x9], 1), 2, False)
        outs = []
        for i, f in enumerate((x15, x18, x21)):
            ch = f.shape[1]
            b = self.conv(f"model.22.cv2.{i}.0", f, 3)
            b = self.conv(f"model.22.cv2.{i}.1", b, 3)
            b = self.conv(f"model.22.cv2.{i}.2", b, 1, act=False, out_f32=True)
            c = self.conv(f"model.22.cv3.{i}.0.0", f, 3, groups=ch)
            c = self.conv(f"model.22.cv3.{i}.0.1", c, 1)
            c = self.conv(f"model.22.cv3.{i}.1.0", c, 3, groups=c.shape[1])
            c = self.conv(f"model.22.cv3.{i}.1.1", c, 1)
            c = self.conv(f"model.22.cv3.{i}.2", c, 1, act=False, out_f32=True)
            outs.append(torch.cat([b, c], 1))
        return outs

    @staticmethod
    def decode(raw: List[torch.Tensor], strides=(8, 16, 32)) -> torch.Tensor:
        """Ultralytics ``Detect`` inference path -> [B, 4+nc, A] (cx,cy,w,h,cls...)."""
        B = raw[0].shape[0]
        no = raw[0].shape[1]
        x_cat = torch.cat([r.reshape(B, no, -1) for r in raw], 2)
        box, cls = x_cat[:, :64], x_cat[:, 64:]
        anchors, svec = [], []
        for r, s in zip(raw, strides):
            h, w = r.shape[2:]
            sx = torch.arange(w, dtype=torch.float32) + 0.5
            sy = torch.arange(h, dtype=torch.float32) + 0.5
            yy, xx = torch.meshgrid(sy, sx, indexing="ij")
            anchors.append(torch.stack((xx, yy), -1).view(-1, 2))
            svec.append(torch.full((h * w, 1), float(s)))
        anchors = torch.cat(anchors).transpose(0, 1)        # [2, A]
        svec = torch.cat(svec).transpose(0, 1)              # [1, A]
        b, _, a = box.shape
        dist = box.view(b, 4, 16, a).transpose(2, 1).softmax(1)   # [B,16,4,A]
        dist = (dist * torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)).sum(1)  # [B,4,A]
        lt, rb = dist.chunk(2, 1)
        x1y1 = anchors.unsqueeze(0) - lt
        x2y2 = anchors.unsqueeze(0) + rb
        c_xy = (x1y1 + x2y2) / 2
        wh = x2y2 - x1y1
        dbox = torch.cat((c_xy, wh), 1) * svec
        return torch.cat((dbox, cls.sigmoid()), 1)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.decode(self.raw_head(x))


V7_ANCHORS = ((12, 16, 19, 36, 40, 28), (36, 75, 76, 55, 72, 146), (142, 110, 192, 243, 459, 401))


class YoloV7Oracle(_Net):
    """Canonical yolov7 deploy graph, module-indexed names."""

    def c(self, i, x, k=1, s=1):
        return self.conv(f"model.{i}", x, k, s)

    def elan(self, i, x):
        a = self.c(i, x); b = self.c(i + 1, x)
        b1 = self.c(i + 2, b, 3); b2 = self.c(i + 3, b1, 3)
        b3 = self.c(i + 4, b2, 3); b4 = self.c(i + 5, b3, 3)
        return self.c(i + 7, torch.cat([b4, b2, b, a], 1))

    def elan_h(self, i, x):
        a = self.c(i, x); b = self.c(i + 1, x)
        b1 = self.c(i + 2, b, 3); b2 = self.c(i + 3, b1, 3)
        b3 = self.c(i + 4, b2, 3); b4 = self.c(i + 5, b3, 3)
        return self.c(i + 7, torch.cat([b4, b3, b2, b1, b, a], 1))

    def mp(self, i, x):
        p = self.c(i + 1, F.max_pool2d(x, 2, 2))
        q = self.c(i + 3, self.c(i + 2, x), 3, 2)
        return torch.cat([q, p], 1)

    def sppcspc(self, x):
        n = "model.51"
        x1 = self.conv(f"{n}.cv4", self.conv(f"{n}.cv3", self.conv(f"{n}.cv1", x), 3))
        y1 = self.conv(f"{n}.cv6", self.conv(f"{n}.cv5", torch.cat(
            [x1] + [F.max_pool2d(x1, k, 1, k // 2) for k in (5, 9, 13)], 1)), 3)
        y2 = self.conv(f"{n}.cv2", x)
        return self.conv(f"{n}.cv7", torch.cat([y1, y2], 1))

    @torch.no_grad()
    def raw_head(self, x):
        x = self.c(0, x, 3, 1); x = self.c(1, x, 3, 2); x = self.c(2, x, 3, 1); x = self.c(3, x, 3, 2)
        x11 = self.elan(4, x)
        p3 = self.elan(17, self.mp(12, x11))
        p4 = self.elan(30, self.mp(25, p3))
        p5 = self.elan(43, self.mp(38, p4))
        n51 = self.sppcspc(p5)
        up = F.interpolate(self.c(52, n51), scale_factor=2, mode="nearest")
        n63 = self.elan_h(56, torch.cat([self.c(54, p4), up], 1))
        up = F.interpolate(self.c(64, n63), scale_factor=2, mode="nearest")
        n75 = self.elan_h(68, torch.cat([self.c(66, p3), up], 1))
        n88 = self.elan_h(81, torch.cat([self.mp(76, n75), n63], 1))
        n101 = self.elan_h(94, torch.cat([self.mp(89, n88), n51], 1))
        outs = []
        for i, f in enumerate((n75, n88, n101)):
            r = self.c(102 + i, f, 3)
            outs.append(self.conv(f"model.105.m.{i}", r, 1, act=False, out_f32=True))
        return outs

    @staticmethod
    def decode(raw, strides=(8, 16, 32)):
        """-> [B, 25200, 5+nc] rows (cx,cy,w,h,obj,cls...) in input pixels."""
        z = []
        for r, s, anc in zip(raw, strides, V7_ANCHORS):
            B, ch, h, w = r.shape
            no = ch // 3
            y = r.view(B, 3, no, h, w).permute(0, 1, 3, 4, 2).sigmoid()
            yv, xv = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
            grid = torch.stack((xv, yv), -1).view(1, 1, h, w, 2)
            ag = torch.tensor(anc, dtype=torch.float32).view(1, 3, 1, 1, 2)
            xy = (y[..., 0:2] * 2.0 - 0.5 + grid) * float(s)
            wh = (y[..., 2:4] * 2.0) ** 2 * ag
            z.append(torch.cat((xy, wh, y[..., 4:]), -1).view(B, -1, no))
        return torch.cat(z, 1)

    @torch.no_grad()
    def forward(self, x):
        return self.decode(self.raw_head(x))


def make_oracle(arch: str, weights, emulate_bf16=False):
    if arch.startswith("yolov8") or arch == "v8":
        return YoloV8mOracle(weights, emulate_bf16)
    return YoloV7Oracle(weights, emulate_bf16)
