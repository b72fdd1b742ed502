"""Synthetic stand-in for the network tile sources (``_script/wms_handler.py``, ``_script/xyz_handler.py``;
out of scope: they are HTTP clients and there is no network).  ``fetch_batch`` keeps the return shape
``GPUHandler.process_batch`` accepts -- the ``XYZHandler`` one, a list of one-element lists
``[(PIL.Image, (lon_min, lat_min, lon_max, lat_max), None)]`` (``xyz_handler.py:170``, ``gpu_handler.py:156-161``) --
and the content of a tile is a deterministic function of its bounding box."""
from __future__ import annotations

import zlib
from typing import List, Sequence, Tuple

import numpy as np

from . import synth


class SyntheticTileHandler:
    def __init__(self, size: int = 640, seed: int = 0, **_ignored):
        self.size, self.seed = int(size), int(seed)
        self.failed_tiles: List[Tuple] = []

    def get_single_image(self, bbox):
        from PIL import Image
        key = zlib.crc32(np.asarray(bbox, dtype=np.float64).round(9).tobytes())
        arr = synth.make_block(self.seed, int(key >> 16), int(key & 0xFFFF), self.size, self.size,
                               cars=max(4, (self.size * self.size * 24) // (512 * 512)))
        return [(Image.fromarray(arr), tuple(bbox), None)]

    def fetch_batch(self, tiles: Sequence, progress_bar=None) -> List:
        out = []
        for bbox in tiles:
            out.append(self.get_single_image(bbox))
            if progress_bar is not None:
                progress_bar.update(1)
        return out
