"""ORACLE (test infrastructure, not product code) -- NumPy restatement of the reference's
test-time-augmentation variants:

* ``_script/gpu_handler.py:94-123``  lighting variants: original, CLAHE(3.0, 8x8) on L of LAB, PIL brightness x2.0, gamma 2.0
* ``_script/gpu_handler.py:125-140`` occlusion variant: CLAHE(4.0, 4x4)
* ``_script/gpu_handler_archive.py:67-122`` archived set: brightness 1.8, (brightness 1.4|1.6 -> contrast 1.3) chain, gamma 1.5,
  CLAHE (2.0, 8x8) / (4.0, 4x4) / (3.0, 16x16)
* ``_script/gpu_handler.py:220-285`` ``_process_tensors`` / ``_get_confidence_adjustment``

The arithmetic lives in OpenCV and Pillow (un-vendored dependencies of the reference: opencv 4.10.0.84, pillow 10.4.0 in
its notebook logs; cv2 4.13 / Pillow 12.2 in this image).  Every function below restates the library's published
algorithm and is PINNED against the library itself in ``tests/test_cpu_tta.py``: the two colour conversions over all
2^24 inputs, CLAHE / brightness / contrast / gamma on the reference's ``test_tile.jpg`` and on random images.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import this module.
"""
from __future__ import annotations

import importlib.util
import os
from typing import List, Sequence, Tuple

import numpy as np

f32 = np.float32

_spec = importlib.util.spec_from_file_location(
    "gen_lab_tables", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools", "gen_lab_tables.py"))
_gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_gen)
_T = None


def lab_tables() -> dict:
    global _T
    if _T is None:
        _T = _gen.tables()
    return _T


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


# ---------------------------------------------------------------------------------
# cv2.cvtColor(img, cv2.COLOR_RGB2LAB) on uint8 -- gpu_handler.py:104, :128
# ---------------------------------------------------------------------------------
def rgb2lab_u8(rgb: np.ndarray) -> np.ndarray:
    t = lab_tables()
    C = t["fwd"]
    R, G, B = (t["srgb_gamma"][rgb[..., k]] for k in range(3))
    fX = t["cbrt"][_descale(R * C[0] + G * C[1] + B * C[2], 12)]
    fY = t["cbrt"][_descale(R * C[3] + G * C[4] + B * C[5], 12)]
    fZ = t["cbrt"][_descale(R * C[6] + G * C[7] + B * C[8], 12)]
    l_scale = (116 * 255 + 50) // 100
    l_shift = -((16 * 255 * (1 << 15) + 50) // 100)
    L = _descale(l_scale * fY + l_shift, 15)
    a = _descale(500 * (fX - fY) + 128 * (1 << 15), 15)
    b = _descale(200 * (fY - fZ) + 128 * (1 << 15), 15)
    return np.clip(np.stack([L, a, b], -1), 0, 255).astype(np.uint8)


# ---------------------------------------------------------------------------------
# cv2.cvtColor(lab, cv2.COLOR_LAB2RGB) on uint8 -- gpu_handler.py:110, :136
# ---------------------------------------------------------------------------------
_MIN_AB = -8145
_BASE = 1 << 14


def _ab_to_xz(v: np.ndarray) -> np.ndarray:
    """f^-1 of CIE Lab in 2^-14 fixed point; C integer division truncates toward zero."""
    def cdiv(a, b):
        return np.sign(a) * (np.abs(a) // b)
    lo = cdiv(v * 108, 841) - _BASE * 16 // 116 * 108 // 841
    hi = cdiv(cdiv(v * v, _BASE) * v, _BASE)
    return np.where(v <= 3390, lo, hi)


def lab2rgb_u8(lab: np.ndarray) -> np.ndarray:
    t = lab_tables()
    C = t["inv"]
    LL, aa, bb = (lab[..., k].astype(np.int64) for k in range(3))
    y, ify = t["l_to_y"][LL], t["l_to_fy"][LL]
    adiv = ((5 * aa * 53687 + (1 << 7)) >> 13) - 128 * _BASE // 500
    bdiv = ((bb * 41943 + (1 << 4)) >> 9) - 128 * _BASE // 200 + 1
    x, z = _ab_to_xz(ify + adiv), _ab_to_xz(ify - bdiv)
    shift = 12 + (14 - 12)
    out = np.stack([_descale(C[0] * x + C[1] * y + C[2] * z, shift), _descale(C[3] * x + C[4] * y + C[5] * z, shift),
                    _descale(C[6] * x + C[7] * y + C[8] * z, shift)], -1)
    return t["inv_gamma"][np.clip(out, 0, 4095)].astype(np.uint8)


# ---------------------------------------------------------------------------------
# cv2.createCLAHE(clipLimit, tileGridSize).apply(l) -- gpu_handler.py:106-107, :131-133
# ---------------------------------------------------------------------------------
def _reflect101(k: np.ndarray, n: int) -> np.ndarray:
    return np.where(k < n, k, 2 * (n - 1) - k)


def clahe_luts(src: np.ndarray, clip: float, tx: int, ty: int) -> Tuple[np.ndarray, int, int]:
    h, w = src.shape
    eh, ew = h, w
    if h % ty or w % tx:            # copyMakeBorder(src, 0, ty - h % ty, 0, tx - w % tx, BORDER_REFLECT_101): when only one
        eh, ew = h + (ty - h % ty), w + (tx - w % tx)      # axis is ragged the other still grows by a whole ty / tx
        src = src[_reflect101(np.arange(eh), h)][:, _reflect101(np.arange(ew), w)]
    th, tw = eh // ty, ew // tx
    area = tw * th
    lut_scale = f32(255) / f32(area)
    limit = max(int(clip * area / 256), 1) if clip > 0 else 0
    luts = np.zeros((ty, tx, 256), np.uint8)
    for j in range(ty):
        for i in range(tx):
            hist = np.bincount(src[j * th:(j + 1) * th, i * tw:(i + 1) * tw].ravel(), minlength=256).astype(np.int64)
            if limit > 0:
                clipped = int(np.maximum(hist - limit, 0).sum())
                hist = np.minimum(hist, limit)
                batch = clipped // 256
                resid = clipped - batch * 256
                hist += batch
                if resid:
                    step = max(256 // resid, 1)
                    hist[np.arange(0, 256, step)[:resid]] += 1
            luts[j, i] = np.clip(np.rint(np.cumsum(hist).astype(f32) * lut_scale), 0, 255).astype(np.uint8)
    return luts, tw, th


def clahe_apply(src: np.ndarray, clip: float, tx: int, ty: int) -> np.ndarray:
    luts, tw, th = clahe_luts(src, clip, tx, ty)
    h, w = src.shape

    def axis(n, inv, tiles):
        f = np.arange(n).astype(f32) * inv - f32(0.5)
        t1 = np.floor(f).astype(np.int64)
        a = (f - t1.astype(f32)).astype(f32)
        return np.maximum(t1, 0), np.minimum(t1 + 1, tiles - 1), a, f32(1) - a
    x1, x2, xa, xa1 = axis(w, f32(1.0) / f32(tw), tx)
    y1, y2, ya, ya1 = axis(h, f32(1.0) / f32(th), ty)
    s = src.astype(np.int64)
    l11 = luts[y1[:, None], x1[None, :], s].astype(f32)
    l12 = luts[y1[:, None], x2[None, :], s].astype(f32)
    l21 = luts[y2[:, None], x1[None, :], s].astype(f32)
    l22 = luts[y2[:, None], x2[None, :], s].astype(f32)
    top = l11 * xa1[None, :] + l12 * xa[None, :]
    bot = l21 * xa1[None, :] + l22 * xa[None, :]
    res = top * ya1[:, None] + bot * ya[:, None]
    return np.clip(np.rint(res), 0, 255).astype(np.uint8)


def clahe_rgb(rgb: np.ndarray, clip: float, tx: int, ty: int) -> np.ndarray:
    """RGB2LAB -> CLAHE on L -> merge -> LAB2RGB (gpu_handler.py:104-110)."""
    lab = rgb2lab_u8(rgb)
    lab[..., 0] = clahe_apply(np.ascontiguousarray(lab[..., 0]), clip, tx, ty)
    return lab2rgb_u8(lab)


# ---------------------------------------------------------------------------------
# PIL ImageEnhance.Brightness / Contrast (Image.blend with a constant image) -- gpu_handler.py:113-115,
# gpu_handler_archive.py:75-84.  ImagingBlend: float32 arithmetic, truncating store.
# ---------------------------------------------------------------------------------
def blend_lut(const: int, factor: float) -> np.ndarray:
    """Image.blend(constant image, image, factor) as a 256-entry table.  Pillow 12 (this image) evaluates
    ``in1 + alpha * (in2 - in1)`` in float32 (separate multiply and add) for interpolation and extrapolation alike,
    clips to [0, 255] and truncates -- checked against Image.blend for every (constant, value) pair in the CPU tests."""
    a = f32(factor)
    v = np.arange(256, dtype=np.int64)
    t = f32(const) + a * (v - const).astype(f32)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, np.clip(t, 0, 255).astype(np.int64))).astype(np.uint8)


def brightness(rgb: np.ndarray, factor: float) -> np.ndarray:
    return blend_lut(0, factor)[rgb]


def grey_mean(rgb: np.ndarray) -> int:
    """``int(ImageStat.Stat(image.convert("L")).mean[0] + 0.5)`` (ImageEnhance.Contrast.__init__)."""
    r, g, b = (rgb[..., k].astype(np.int64) for k in range(3))
    grey = (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16
    return int(int(grey.sum()) / grey.size + 0.5)


def contrast(rgb: np.ndarray, factor: float) -> np.ndarray:
    return blend_lut(grey_mean(rgb), factor)[rgb]


def gamma_lut(gamma: float) -> np.ndarray:
    """``(np.power(img / 255.0, 1.0 / gamma) * 255.0).astype(np.uint8)`` -- gpu_handler.py:118-121."""
    return (np.power(np.arange(256) / 255.0, 1.0 / gamma) * 255.0).astype(np.uint8)


def lighting_variations(rgb: np.ndarray) -> List[np.ndarray]:
    """gpu_handler.py:94-123 (uint8 RGB images; the caller's ``_prepare_tensor`` makes them BGR f32 CHW / 255)."""
    return [rgb, clahe_rgb(rgb, 3.0, 8, 8), brightness(rgb, 2.0), gamma_lut(2.0)[rgb]]


def occlusion_variations(rgb: np.ndarray) -> List[np.ndarray]:
    """gpu_handler.py:125-140."""
    return [clahe_rgb(rgb, 4.0, 4, 4)]


def archive_variations(rgb: np.ndarray) -> List[np.ndarray]:
    """gpu_handler_archive.py:57-122."""
    out = [rgb, brightness(rgb, 1.8)]
    s = rgb
    for b in (1.4, 1.6):
        s = contrast(brightness(s, b), 1.3)
        out.append(s)
    out.append(gamma_lut(1.5)[rgb])
    for clip, t in ((2.0, 8), (4.0, 4), (3.0, 16)):
        out.append(clahe_rgb(rgb, clip, t, t))
    return out


CONF_ADJUST = {0: 1.0, 1: 0.95, 2: 0.90, 3: 0.92, 4: 0.88}      # gpu_handler.py:274-283


def confidence_adjustment(i: int) -> float:
    return CONF_ADJUST.get(i, 0.85)


def process_tensors_rows(rows_per_variant: Sequence[np.ndarray], bbox: Sequence[float], thr: float) -> np.ndarray:
    """``_process_tensors`` for one tile (gpu_handler.py:226-256) from the rows of each variant.

    ``boxes[:, 4] *= adj`` (float32), ``> thr`` (strict), concatenation in variant order, then the float32 CUDA-tensor
    georeferencing: ``centers = boxes[:, :2] / 640`` is ``x * (1.0f / 640.0f)`` on a CUDA tensor (torch multiplies by the
    reciprocal of a scalar divisor), ``lon = bbox[0] + centers_x * lon_offset`` with both Python scalars cast to float32.
    Returns [K, 3] float32 (lon, lat, conf).
    """
    kept = []
    for i, rows in enumerate(rows_per_variant):
        r = rows.astype(np.float32).copy()
        r[:, 4] = r[:, 4] * f32(confidence_adjustment(i))
        kept.append(r[r[:, 4] > f32(thr)])
    r = np.concatenate(kept, 0) if kept else np.zeros((0, 6), np.float32)
    inv = f32(1.0) / f32(640.0)
    cx, cy = r[:, 0] * inv, r[:, 1] * inv
    lon_off, lat_off = f32(bbox[2] - bbox[0]), f32(bbox[3] - bbox[1])
    lon = f32(bbox[0]) + cx * lon_off
    lat = f32(bbox[3]) - cy * lat_off
    return np.stack([lon, lat, r[:, 4]], 1).astype(np.float32)
