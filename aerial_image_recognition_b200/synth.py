"""Synthetic orthophoto tiles and mosaics.

Stand-in for the network tile sources ``_script/wms_handler.py:196-249`` and
``_script/xyz_handler.py:228-248`` (out of scope: they are HTTP clients).  The
generator is procedural in ``(seed, block_y, block_x)`` so any rank can
materialise any part of a mosaic without host traffic, and the same function is
used for single tiles (BASELINE config C2/C3) and the sliding-window mosaic (C4).

Statistics follow the reference's own imagery: per-channel mean ~90-145 and std
~45-55 (``img/srodmiescie.tiff.aux.xml`` band stats; ``test_tile.jpg``
mean 144 / std 48), low-frequency background plus car-sized bright/dark
rectangles (~45x20 px at 10 cm/px, i.e. ``tile_size_meters 64`` / 640 px,
``_script/config.py:13``).
"""
from __future__ import annotations

import numpy as np

BLOCK = 512  # mosaic generation granularity (px)


def _lowfreq(rng, h, w, cells=8):
    gh, gw = h // cells + 2, w // cells + 2
    coarse = rng.normal(0.0, 1.0, size=(gh, gw, 3)).astype(np.float32)
    ys = (np.arange(h, dtype=np.float32) + 0.5) / cells
    xs = (np.arange(w, dtype=np.float32) + 0.5) / cells
    y0 = np.floor(ys).astype(np.int64); x0 = np.floor(xs).astype(np.int64)
    fy = (ys - y0)[:, None, None]; fx = (xs - x0)[None, :, None]
    a = coarse[y0][:, x0]; b = coarse[y0][:, x0 + 1]
    c = coarse[y0 + 1][:, x0]; d = coarse[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def make_block(seed: int, by: int, bx: int, h: int = BLOCK, w: int = BLOCK, cars: int = 24) -> np.ndarray:
    """uint8 [h, w, 3] block, deterministic in (seed, by, bx)."""
    rng = np.random.default_rng([1234, seed, by, bx])
    base = np.array([118.0, 121.0, 112.0], dtype=np.float32)
    img = base + 38.0 * _lowfreq(rng, h, w, 64) + 22.0 * _lowfreq(rng, h, w, 8)
    img += rng.normal(0.0, 9.0, size=(h, w, 3)).astype(np.float32)
    for _ in range(cars):
        cw, ch = (45, 20) if rng.random() < 0.5 else (20, 45)
        cw += int(rng.integers(-4, 5)); ch += int(rng.integers(-3, 4))
        x = int(rng.integers(0, max(1, w - cw))); y = int(rng.integers(0, max(1, h - ch)))
        col = rng.choice([25.0, 60.0, 200.0, 235.0]) + rng.normal(0, 8.0, size=3)
        img[y:y + ch, x:x + cw] = col.astype(np.float32)
        img[y + ch // 4:y + 3 * ch // 4, x + cw // 4:x + 3 * cw // 4] *= 0.8
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def make_tiles(n: int, size: int = 640, seed: int = 0) -> np.ndarray:
    """uint8 [n, size, size, 3] independent tiles (BASELINE configs C2 / C3)."""
    out = np.empty((n, size, size, 3), dtype=np.uint8)
    for i in range(n):
        out[i] = make_block(seed, 1_000_000 + i, 0, size, size, cars=max(4, (size * size * 24) // (512 * 512)))
    return out


def make_mosaic(height: int, width: int, seed: int = 0, y0: int = 0, y1: int | None = None) -> np.ndarray:
    """uint8 rows [y0, y1) of a ``height x width`` mosaic assembled from 512-px blocks."""
    y1 = height if y1 is None else y1
    out = np.empty((y1 - y0, width, 3), dtype=np.uint8)
    for by in range(y0 // BLOCK, (y1 + BLOCK - 1) // BLOCK):
        ys, ye = max(y0, by * BLOCK), min(y1, (by + 1) * BLOCK)
        for bx in range((width + BLOCK - 1) // BLOCK):
            xs, xe = bx * BLOCK, min(width, (bx + 1) * BLOCK)
            blk = make_block(seed, by, bx)
            out[ys - y0:ye - y0, xs:xe] = blk[ys - by * BLOCK:ye - by * BLOCK, :xe - xs]
    return out


MOSAIC_POOL = 32   # distinct procedural blocks a device-side mosaic is assembled from


def mosaic_block_pool(seed: int = 77) -> np.ndarray:
    """uint8 [MOSAIC_POOL, 512, 512, 3]: the blocks ``mosaic_band_device`` tiles a mosaic with."""
    return np.stack([make_block(seed, 0, i) for i in range(MOSAIC_POOL)])


def mosaic_band_device(pool, height: int, width: int, y_lo: int, y_hi: int, seed: int = 5):
    """uint8 CUDA tensor [y_hi - y_lo, width, 3]: pixel rows [y_lo, y_hi) of the ``height x width`` mosaic whose 512-px
    block (by, bx) is ``pool[hash(seed, by, bx)]`` (``pool`` = ``mosaic_block_pool`` on the device).  The content depends on the
    *global* block coordinates only, so every rank of a tile-sharded run (BASELINE config C4) materialises its own band on
    its own device with no host traffic, and the mosaic is the same for every world size."""
    import torch
    by0, by1 = y_lo // BLOCK, (y_hi + BLOCK - 1) // BLOCK
    nbx = (width + BLOCK - 1) // BLOCK
    by = np.arange(by0, by1, dtype=np.uint64)[:, None]
    bx = np.arange(nbx, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        h = (by * np.uint64(0x9E3779B97F4A7C15) + bx * np.uint64(0xC2B2AE3D27D4EB4F) + np.uint64(seed) * np.uint64(0x165667B19E3779F9))
    h ^= h >> np.uint64(29)
    idx = torch.from_numpy((h % np.uint64(pool.shape[0])).astype(np.int64)).to(pool.device)
    rows = []
    for r in range(by1 - by0):                                  # one block row at a time keeps the temporary small
        strip = pool[idx[r]].permute(1, 0, 2, 3).reshape(BLOCK, nbx * BLOCK, 3)[:, :width]
        lo = max(y_lo, (by0 + r) * BLOCK) - (by0 + r) * BLOCK
        hi = min(y_hi, (by0 + r + 1) * BLOCK) - (by0 + r) * BLOCK
        rows.append(strip[lo:hi])
    return torch.cat(rows).contiguous()
