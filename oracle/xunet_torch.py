"""ORACLE (test infrastructure, not product code) -- PyTorch fp32 CPU statement of the
segmentation stand-in for BASELINE config C5 ("ramp XUnet 256").

PARITY UNPINNED, and more so than for the detectors: the reference holds only the blob's NAME
(``.MISSING_LARGE_BLOBS:3``) -- no code, no architecture, no call site.  What is
stated here is therefore not the reference's algorithm but the *declared stand-in* of
``aerial_image_recognition_b200/graph.py::build_xunet`` ([EXT] ramp: EfficientNet-B0-encoder U-Net,
256 x 256 x 3 in, 4-class softmax out), written the conventional way -- NCHW modules, ``torch.cat`` for
the skips -- so that it is independent of the engine's offset-write graph.  It pins the engine's
wiring and arithmetic for this graph, nothing about ramp's weights or accuracy.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from .yolo_torch import _Net

# EfficientNet-B0 stage table with the widths of stages 2-5 rounded up to multiples of 32 (expand, channels, repeats, stride) and the
# segmentation_models U-Net decoder widths; see build_xunet for the differences from B0 (3x3 depthwise everywhere, stride 2 =
# stride-1 depthwise + 2x2 max-pool, no squeeze-and-excitation, SiLU in the decoder).
STAGES = ((1, 16, 1, 1), (6, 32, 2, 2), (6, 64, 2, 2), (6, 96, 3, 2), (6, 128, 3, 1), (6, 192, 4, 2), (6, 320, 1, 1))
DECODER = (256, 128, 64, 32, 16)
SKIP_STAGES = (1, 2, 3, 5)          # stages whose last block feeds a decoder level


class XUnetOracle(_Net):
    def __init__(self, weights: Dict[str, np.ndarray], emulate_bf16=False, nc: int = 4):
        super().__init__(weights, emulate_bf16)
        self.nc = nc

    def mbconv(self, name, x, expand, cout, stride):
        cin = x.shape[1]
        t = x if expand == 1 else self.conv(f"{name}.expand", x, 1, 1)
        t = self.conv(f"{name}.dw", t, 3, 1, groups=t.shape[1])
        if stride == 2:
            t = F.max_pool2d(t, 2, 2)
        return self.conv(f"{name}.project", t, 1, 1, act=False, res=x if (stride == 1 and cin == cout) else None)

    def logits(self, x: torch.Tensor) -> torch.Tensor:
        """x: float32 [B, 3, H, W] in [0, 1] -> logits [B, nc, H, W]."""
        x = self.conv("encoder.stem", x, 3, 2)
        skips = []
        for si, (e, c, r, s) in enumerate(STAGES, start=1):
            for bi in range(r):
                x = self.mbconv(f"encoder.s{si}.b{bi}", x, e, c, s if bi == 0 else 1)
            if si in SKIP_STAGES:
                skips.append(x)
        for lvl, cdec in enumerate(DECODER):
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            if lvl < 4:
                x = torch.cat([x, skips[3 - lvl]], 1)
            x = self.conv(f"decoder.b{lvl}.conv1", x, 3, 1)
            x = self.conv(f"decoder.b{lvl}.conv2", x, 3, 1)
        return self.conv("segmentation_head", x, 3, 1, act=False, out_f32=True)

    def forward(self, x: torch.Tensor):
        """labels uint8 [B, H, W] (argmax, first maximum wins) and the softmax probability of that class."""
        z = self.logits(x)
        return z.argmax(1).to(torch.uint8), torch.softmax(z, 1).amax(1)
