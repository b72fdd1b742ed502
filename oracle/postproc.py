"""ORACLE (test infrastructure, not product code) -- NumPy / torchvision restatement
of everything the reference does after ``session.run``.

PARITY UNPINNED: the reference has no tests or golden vectors for these steps
(SURVEY.md section 4); each function cites the reference lines it follows and is
checked in ``tests/`` against hand-computed cases.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import this module.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np


# ---------------------------------------------------------------------------------
# confidence filter -- simple_detector.py:479-481, :672-673; gpu_handler.py:166-173
# ---------------------------------------------------------------------------------
def filter_rows(rows: np.ndarray, thr: float, inclusive: bool = True) -> np.ndarray:
    """``boxes[boxes[:, 4] >= thr]`` (column 4 only; class scores ignored)."""
    m = rows[:, 4] >= np.float32(thr) if inclusive else rows[:, 4] > np.float32(thr)
    return rows[m]


def top_k_rows(rows: np.ndarray, k: int = 10) -> np.ndarray:
    """``np.argsort(-conf)[:10]`` -- gpu_handler.py:173 (debug leftover, kept as a flag)."""
    return rows[np.argsort(-rows[:, 4], kind="stable")[:k]]


# ---------------------------------------------------------------------------------
# Ultralytics non_max_suppression [EXT], the post-processing behind ``model(window)`` at
# x_arch/02_analyze_images:1 (cell 6); args conf=0.25 iou=0.7 max_det=300
# (x_arch/01_train_tokyo.ipynb:1 (cell 15 output)).
# ---------------------------------------------------------------------------------
def ultralytics_nms(pred: np.ndarray, conf_thres=0.25, iou_thres=0.7, max_det=300,
                    max_nms=30000, max_wh=7680, agnostic=False) -> List[np.ndarray]:
    """pred: [B, 4+nc, A] (cx,cy,w,h,cls...).  Returns per image [n,6] (x1,y1,x2,y2,conf,cls)."""
    import torch
    import torchvision
    p = torch.from_numpy(np.ascontiguousarray(pred)).float()
    nc = p.shape[1] - 4
    xc = p[:, 4:4 + nc].amax(1) > conf_thres
    p = p.transpose(-1, -2).clone()
    xy, wh = p[..., :2].clone(), p[..., 2:4].clone()
    p[..., :2] = xy - wh / 2
    p[..., 2:4] = xy + wh / 2
    out = []
    for xi, x in enumerate(p):
        x = x[xc[xi]]
        if not x.shape[0]:
            out.append(np.zeros((0, 6), np.float32)); continue
        box, cls = x[:, :4], x[:, 4:4 + nc]
        conf, j = cls.max(1, keepdim=True)
        x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]
        n = x.shape[0]
        if not n:
            out.append(np.zeros((0, 6), np.float32)); continue
        if n > max_nms:
            x = x[x[:, 4].argsort(descending=True)[:max_nms]]
        c = x[:, 5:6] * (0 if agnostic else max_wh)
        scores = x[:, 4]
        boxes = x[:, :4] + c
        i = torchvision.ops.nms(boxes, scores, iou_thres)[:max_det]
        out.append(x[i].numpy())
    return out


def nms_greedy_reference(boxes: np.ndarray, scores: np.ndarray, iou_thres: float) -> np.ndarray:
    """Plain-loop greedy NMS with torchvision's arithmetic (fp32, suppress iff IoU > thr,
    ties broken by lower index).  Cross-check for ``torchvision.ops.nms``."""
    order = np.lexsort((np.arange(len(scores)), -scores.astype(np.float32)))
    b = boxes.astype(np.float32)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    keep, dead = [], np.zeros(len(b), bool)
    for ii, i in enumerate(order):
        if dead[i]:
            continue
        keep.append(i)
        rest = order[ii + 1:]
        xx1 = np.maximum(b[i, 0], b[rest, 0]); yy1 = np.maximum(b[i, 1], b[rest, 1])
        xx2 = np.minimum(b[i, 2], b[rest, 2]); yy2 = np.minimum(b[i, 3], b[rest, 3])
        w = np.maximum(np.float32(0), xx2 - xx1); h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = inter / (area[i] + area[rest] - inter)
        dead[rest[iou > np.float32(iou_thres)]] = True
    return np.asarray(keep, dtype=np.int64)


def v8_rows_adapter(pred_b: np.ndarray) -> np.ndarray:
    """[4+nc, A] -> [A, 6] rows (cx,cy,w,h,conf=max cls,cls) so that the reference's
    ``rows[:, 4] >= thr`` keeps its meaning for the v8 head (an adapter)."""
    cls = pred_b[4:]
    return np.concatenate([pred_b[:4].T, cls.max(0)[:, None], cls.argmax(0)[:, None].astype(np.float32)], 1)


def scale_boxes(boxes: np.ndarray, h0: int, w0: int, size: int = 640) -> np.ndarray:
    """Ultralytics ``scale_boxes`` [EXT]: undo the letterbox, clip to the source window."""
    gain = min(size / h0, size / w0)
    padx = round((size - w0 * gain) / 2 - 0.1)
    pady = round((size - h0 * gain) / 2 - 0.1)
    b = boxes.astype(np.float32).copy()
    b[:, [0, 2]] -= np.float32(padx)
    b[:, [1, 3]] -= np.float32(pady)
    b[:, :4] /= np.float32(gain)
    b[:, [0, 2]] = b[:, [0, 2]].clip(0, w0)
    b[:, [1, 3]] = b[:, [1, 3]].clip(0, h0)
    return b


# ---------------------------------------------------------------------------------
# georeferencing
# ---------------------------------------------------------------------------------
def georef_bounds(x: float, y: float, west: float, east: float, south: float, north: float,
                  model_size: int = 640, crop_size: float = 864):
    """simple_detector.py:487-494 / :519-528.  Evaluated in float64 as under the
    reference's NumPy 1.26: cast the f32 box centre to a
    Python float first."""
    x = float(x); y = float(y)
    x_frac = x / model_size
    y_frac = y / model_size
    x_img = x_frac * crop_size
    y_img = y_frac * crop_size
    lon = west + x_frac * (east - west)
    lat = north - y_frac * (north - south)
    return lon, lat, x_img, y_img


def georef_gpuhandler(x: float, y: float, lon_min, lat_min, lon_max, lat_max):
    """_script/gpu_handler.py:182-190: keeps the ``*864/864`` round trip."""
    x = float(x); y = float(y)
    x_864 = (x / 640) * 864
    y_864 = (y / 640) * 864
    lon = lon_min + (x_864 / 864) * (lon_max - lon_min)
    lat = lat_max - (y_864 / 864) * (lat_max - lat_min)
    return lon, lat


def georef_affine(px: float, py: float, gt: Sequence[float]):
    """``pixel_to_geo`` at x_arch/02_analyze_images:1 (cell 6): GDAL 6-term affine."""
    px = float(px); py = float(py)
    gx = gt[0] + px * gt[1] + py * gt[2]
    gy = gt[3] + px * gt[4] + py * gt[5]
    return gx, gy


# ---------------------------------------------------------------------------------
# centre-distance greedy dedup -- simple_detector.py:540-596 (primary, inclusive <=)
# and _script/utils.py:212-274 (variant, strict <)
# ---------------------------------------------------------------------------------
def dedup_greedy(x: np.ndarray, y: np.ndarray, conf: np.ndarray, thr: float, inclusive: bool = True) -> np.ndarray:
    """Returns indices (into the input) of the kept detections, in descending-confidence
    order with input order among equal confidences (Python's stable ``sort(reverse=True)``,
    simple_detector.py:565).  Uses a uniform grid instead of the rtree; the accept test is
    the reference's exact ``dx*dx + dy*dy <= thr*thr`` in float64 (:585-587)."""
    n = len(x)
    if n == 0:
        return np.zeros(0, np.int64)
    x = np.asarray(x, np.float64); y = np.asarray(y, np.float64)
    order = sorted(range(n), key=lambda i: conf[i], reverse=True)
    cell = thr if thr > 0 else 1.0
    grid = {}
    keep = []
    t2 = thr * thr
    for i in order:
        cx, cy = int(math.floor(x[i] / cell)), int(math.floor(y[i] / cell))
        dup = False
        for gx in (cx - 1, cx, cx + 1):
            for gy in (cy - 1, cy, cy + 1):
                for j in grid.get((gx, gy), ()):
                    dx = x[i] - x[j]; dy = y[i] - y[j]
                    d2 = dx * dx + dy * dy
                    if (d2 <= t2) if inclusive else (d2 < t2):
                        dup = True; break
                if dup: break
            if dup: break
        if not dup:
            keep.append(i)
            grid.setdefault((cx, cy), []).append(i)
    return np.asarray(keep, np.int64)


def dedup_bruteforce(x, y, conf, thr, inclusive=True) -> np.ndarray:
    """O(n^2) literal restatement (no spatial index) for small cross-checks."""
    n = len(x)
    order = sorted(range(n), key=lambda i: conf[i], reverse=True)
    keep = []
    t2 = thr * thr
    for i in order:
        ok = True
        for j in keep:
            dx = float(x[i]) - float(x[j]); dy = float(y[i]) - float(y[j])
            d2 = dx * dx + dy * dy
            if (d2 <= t2) if inclusive else (d2 < t2):
                ok = False; break
        if ok:
            keep.append(i)
    return np.asarray(keep, np.int64)


# ---------------------------------------------------------------------------------
# WGS84 -> UTM (replaces pyproj, simple_detector.py:546-556).  Krueger series to n^4,
# the published extended transverse-Mercator formulation [EXT]; cannot be bit-checked
# against PROJ here -- sub-millimetre agreement with published test points.
# ---------------------------------------------------------------------------------
_A = 6378137.0
_F = 1 / 298.257223563
_N = _F / (2 - _F)
_AA = _A / (1 + _N) * (1 + _N ** 2 / 4 + _N ** 4 / 64)
_ALPHA = (
    _N / 2 - 2 * _N ** 2 / 3 + 5 * _N ** 3 / 16 + 41 * _N ** 4 / 180,
    13 * _N ** 2 / 48 - 3 * _N ** 3 / 5 + 557 * _N ** 4 / 1440,
    61 * _N ** 3 / 240 - 103 * _N ** 4 / 140,
    49561 * _N ** 4 / 161280,
)


def utm_zone(lon: float, lat: float) -> Tuple[int, bool]:
    """simple_detector.py:546-548: zone from the *first* detection."""
    return int((lon + 180) / 6) + 1, lat > 0


def utm_forward(lon, lat, zone: int, north: bool):
    lon = np.asarray(lon, np.float64); lat = np.asarray(lat, np.float64)
    lon0 = math.radians((zone - 1) * 6 - 180 + 3)
    phi = np.radians(lat); lam = np.radians(lon) - lon0
    e = math.sqrt(_F * (2 - _F))
    s = np.sin(phi)
    t = np.sinh(np.arctanh(s) - e * np.arctanh(e * s))
    xi = np.arctan2(t, np.cos(lam))
    eta = np.arctanh(np.sin(lam) / np.sqrt(1 + t * t))
    x = eta.copy(); y = xi.copy()
    for j, a in enumerate(_ALPHA, start=1):
        x = x + a * np.cos(2 * j * xi) * np.sinh(2 * j * eta)
        y = y + a * np.sin(2 * j * xi) * np.cosh(2 * j * eta)
    k0 = 0.9996
    E = 500000.0 + k0 * _AA * x
    N = k0 * _AA * y + (0.0 if north else 10000000.0)
    return E, N


# ---------------------------------------------------------------------------------
# tile indexing
# ---------------------------------------------------------------------------------
def sliding_windows(h: int, w: int, win: int, stride: int) -> List[Tuple[int, int, int, int]]:
    """x_arch/02_analyze_images:1 (cell 6): ``for y in range(0,h,stride): for x in
    range(0,w,stride)``, window clipped with ``min``.  Returns (x0, y0, x1, y1)."""
    out = []
    for y in range(0, h, stride):
        for x in range(0, w, stride):
            out.append((x, y, min(x + win, w), min(y + win, h)))
    return out


def generate_tiles_metric(minx: float, miny: float, maxx: float, maxy: float, size: float, overlap: float):
    """_script/utils.py:45-63 in an already-metric CRS: y outer, x inner, unclipped,
    step accumulated by repeated addition."""
    tiles = []
    y = miny
    while y < maxy:
        x = minx
        while x < maxx:
            tiles.append((x, y, x + size, y + size))
            x += size * (1 - overlap)
        y += size * (1 - overlap)
    return tiles
