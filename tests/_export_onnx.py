"""Test helper: a conventional ``torch.nn.Module`` YOLOv8m (Ultralytics 8.3.x module tree and names) exported to ``.onnx`` by
torch's own TorchScript ONNX exporter -- the exporter Ultralytics' ``model.export(format="onnx")`` drives -- so that
``onnx_reader`` is exercised on a real exporter's file (Conv -> Sigmoid -> Mul, Split, Concat, MaxPool, Resize, Softmax ...),
not only on files written by its own minimal writer.  The reference loads such a file at ``_script/gpu_handler.py:39-65``.

The ``onnx`` Python package is absent; the exporter's C++ core serialises the ModelProto itself and only a post-processing
step for onnxscript functions imports ``onnx`` -- that step is replaced by the identity here.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class Conv(nn.Module):
    """Ultralytics ``Conv`` in deploy form (BatchNorm folded): Conv2d(bias) + SiLU written as x * sigmoid(x)."""

    def __init__(self, c1, c2, k=1, s=1, g=1, bn=False):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2, groups=g, bias=not bn)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3) if bn else None

    def forward(self, x):
        y = self.conv(x)
        if self.bn is not None:
            y = self.bn(y)
        return y * torch.sigmoid(y)


class Bottleneck(nn.Module):
    def __init__(self, c, shortcut):
        super().__init__()
        self.cv1, self.cv2, self.add = Conv(c, c, 3), Conv(c, c, 3), shortcut

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class C2f(nn.Module):
    def __init__(self, c1, c2, n, shortcut):
        super().__init__()
        self.c = c2 // 2
        self.cv1 = Conv(c1, 2 * self.c, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, shortcut) for _ in range(n))

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        y.extend(m(y[-1]) for m in self.m)
        return self.cv2(torch.cat(y, 1))


class SPPF(nn.Module):
    def __init__(self, c1, c2):
        super().__init__()
        self.cv1, self.cv2 = Conv(c1, c1 // 2, 1), Conv(c1 * 2, c2, 1)

    def forward(self, x):
        y = [self.cv1(x)]
        y.extend(F.max_pool2d(y[-1], 5, 1, 2) for _ in range(3))
        return self.cv2(torch.cat(y, 1))


class DFL(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(16, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)

    def forward(self, x):
        b, _, a = x.shape
        return self.conv(x.view(b, 4, 16, a).transpose(2, 1).softmax(1)).view(b, 4, a)


class Detect(nn.Module):
    def __init__(self, nc, ch):
        super().__init__()
        self.nc, c2, c3 = nc, 64, 192
        self.cv2 = nn.ModuleList(nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 64, 1)) for x in ch)
        self.cv3 = nn.ModuleList(nn.Sequential(nn.Sequential(Conv(x, x, 3, g=x), Conv(x, c3, 1)),
                                               nn.Sequential(Conv(c3, c3, 3, g=c3), Conv(c3, c3, 1)), nn.Conv2d(c3, nc, 1)) for x in ch)
        self.dfl = DFL()

    def forward(self, xs):
        outs = [torch.cat((self.cv2[i](x), self.cv3[i](x)), 1) for i, x in enumerate(xs)]
        b = outs[0].shape[0]
        cat = torch.cat([o.view(b, 64 + self.nc, -1) for o in outs], 2)
        box, cls = cat[:, :64], cat[:, 64:]
        anchors, strides = [], []
        for o, s in zip(outs, (8, 16, 32)):
            h, w = o.shape[2:]
            yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32) + 0.5, torch.arange(w, dtype=torch.float32) + 0.5, indexing="ij")
            anchors.append(torch.stack((xx, yy), -1).view(-1, 2))
            strides.append(torch.full((h * w, 1), float(s)))
        anchors = torch.cat(anchors).transpose(0, 1).unsqueeze(0).to(box.dtype)
        strides = torch.cat(strides).transpose(0, 1).to(box.dtype)
        lt, rb = self.dfl(box).chunk(2, 1)
        x1y1, x2y2 = anchors - lt, anchors + rb
        return torch.cat((torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * strides, cls.sigmoid()), 1)


class YoloV8m(nn.Module):
    """``model.N`` = the layer table of x_arch/01_train_tokyo.ipynb:1 (cell 15 output)."""

    def __init__(self, nc=2):
        super().__init__()
        up = nn.Upsample(scale_factor=2, mode="nearest")
        self.model = nn.ModuleList([
            Conv(3, 48, 3, 2), Conv(48, 96, 3, 2), C2f(96, 96, 2, True), Conv(96, 192, 3, 2), C2f(192, 192, 4, True),
            Conv(192, 384, 3, 2), C2f(384, 384, 4, True), Conv(384, 576, 3, 2), C2f(576, 576, 2, True), SPPF(576, 576),
            up, nn.Identity(), C2f(960, 384, 2, False), up, nn.Identity(), C2f(576, 192, 2, False),
            Conv(192, 192, 3, 2), nn.Identity(), C2f(576, 384, 2, False), Conv(384, 384, 3, 2), nn.Identity(), C2f(960, 576, 2, False),
            Detect(nc, (192, 384, 576))])

    def forward(self, x):
        m = self.model
        x4 = m[4](m[3](m[2](m[1](m[0](x)))))
        x6 = m[6](m[5](x4))
        x9 = m[9](m[8](m[7](x6)))
        x12 = m[12](torch.cat([m[10](x9), x6], 1))
        x15 = m[15](torch.cat([m[13](x12), x4], 1))
        x18 = m[18](torch.cat([m[16](x15), x12], 1))
        x21 = m[21](torch.cat([m[19](x18), x9], 1))
        return m[22]([x15, x18, x21])


def load_deploy_weights(model: nn.Module, w) -> None:
    """``{engine conv name + '.weight' / '.bias'}`` into the module tree (``Conv`` modules hold their Conv2d as ``.conv``)."""
    sd = model.state_dict()
    for k, v in w.items():
        base, kind = k.rsplit(".", 1)
        name = f"{base}.conv.{kind}" if f"{base}.conv.{kind}" in sd else k
        assert name in sd and tuple(sd[name].shape) == tuple(v.shape), (k, name)
        sd[name].copy_(torch.from_numpy(np.ascontiguousarray(v)))


def export(model: nn.Module, path: str, imgsz: int = 64, half: bool = False, train_mode: bool = False) -> None:
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils
    import warnings
    keep = onnx_proto_utils._add_onnxscript_fn
    onnx_proto_utils._add_onnxscript_fn = lambda proto, custom_opsets: proto      # needs the absent `onnx` package; no onnxscript functions here
    try:
        x = torch.zeros(1, 3, imgsz, imgsz)
        if half:
            model, x = model.half(), x.half()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            kw = dict(training=torch.onnx.TrainingMode.TRAINING, do_constant_folding=False) if train_mode else {}
            torch.onnx.export(model if train_mode else model.eval(), x, path, opset_version=13, input_names=["images"], output_names=["output0"],
                              dynamo=False, **kw)
    finally:
        onnx_proto_utils._add_onnxscript_fn = keep
