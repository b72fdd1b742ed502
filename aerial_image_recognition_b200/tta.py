"""Host side of the test-time-augmentation variants (``_script/gpu_handler.py:94-140``, ``:220-285``;
``_script/gpu_handler_archive.py:57-122``, ``:229-246``).

The pixel work runs on the device (``csrc/tta.cu``); what is left for the host is what the reference also does on the
host with scalars: the 256-entry curves of the per-byte variants and the per-view confidence weights.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

# view index -> confidence weight (gpu_handler.py:274-283); anything else 0.85 (:283)
CONFIDENCE_ADJUSTMENTS = {0: 1.0, 1: 0.95, 2: 0.90, 3: 0.92, 4: 0.88}
# archived weights (gpu_handler_archive.py:229-246): views 0-4 -> 1.0, 5-7 -> 0.98, 8-11 -> 0.95, else 0.85
ARCHIVE_ADJUSTMENTS = {**{i: 1.0 for i in range(5)}, 5: 0.98, 6: 0.98, 7: 0.98, 8: 0.95, 9: 0.95, 10: 0.95, 11: 0.95}

# (kind, args) of each view; "clahe": (clip limit, grid), "brightness"/"contrast": factor, "gamma": gamma
LIGHTING_VIEWS: List[Tuple[str, tuple]] = [("original", ()), ("clahe", (3.0, 8)), ("brightness", (2.0,)), ("gamma", (2.0,))]
OCCLUSION_VIEWS: List[Tuple[str, tuple]] = [("clahe", (4.0, 4))]
# archived set: the two chained views are (brightness b -> contrast 1.3) applied cumulatively (gpu_handler_archive.py:79-84)
ARCHIVE_VIEWS: List[Tuple[str, tuple]] = [("original", ()), ("brightness", (1.8,)), ("chain", (1.4, 1.3)), ("chain", (1.6, 1.3)),
                                          ("gamma", (1.5,)), ("clahe", (2.0, 8)), ("clahe", (4.0, 4)), ("clahe", (3.0, 16))]


def confidence_adjustment(variation_index: int, total_variations: int = 5, archive: bool = False) -> float:
    table = ARCHIVE_ADJUSTMENTS if archive else CONFIDENCE_ADJUSTMENTS
    return table.get(variation_index, 0.85)


def brightness_lut(factor: float) -> np.ndarray:
    """``ImageEnhance.Brightness(img).enhance(factor)`` = ``Image.blend(black, img, factor)`` as a byte curve: Pillow's
    ImagingBlend computes ``in1 + alpha * (in2 - in1)`` in float32 with in1 = 0, clips to [0, 255] and truncates."""
    t = np.float32(factor) * np.arange(256, dtype=np.float32)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, np.clip(t, 0, 255).astype(np.int64))).astype(np.uint8)


def gamma_lut(gamma: float) -> np.ndarray:
    """``(np.power(img / 255.0, 1.0 / gamma) * 255.0).astype(np.uint8)`` (gpu_handler.py:118-121) on the 256 byte values."""
    return (np.power(np.arange(256) / 255.0, 1.0 / gamma) * 255.0).astype(np.uint8)
