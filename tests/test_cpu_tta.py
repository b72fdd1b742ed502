"""CPU tests (-m "not gpu") of the test-time-augmentation oracle (oracle/tta.py) against the libraries whose arithmetic
the reference calls: OpenCV (RGB<->Lab, CLAHE) and Pillow (ImageEnhance), SURVEY.md section 8f-4.  These pin the oracle;
tests/test_gpu_tta.py then holds the CUDA kernels to the oracle and to the libraries again."""
import os
import subprocess
import sys

import cv2
import numpy as np
import pytest
from PIL import Image, ImageEnhance

from oracle import tta as OT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tile():
    return np.array(Image.open(os.path.join(ROOT, "tests", "golden", "test_tile_864.png")).convert("RGB"))


def _cube(first: int) -> np.ndarray:
    g = np.arange(256, dtype=np.uint8)
    c = np.zeros((256, 256, 3), np.uint8)
    c[..., 0] = first
    c[..., 1] = g[:, None]
    c[..., 2] = g[None, :]
    return c


def test_rgb2lab_equals_cv2_over_all_inputs():
    bad = 0
    for r in range(256):
        c = _cube(r)
        bad += int((cv2.cvtColor(c, cv2.COLOR_RGB2LAB) != OT.rgb2lab_u8(c)).sum())
    assert bad == 0


def test_lab2rgb_equals_cv2_over_all_inputs():
    bad = 0
    for L in range(256):
        c = _cube(L)
        bad += int((cv2.cvtColor(c, cv2.COLOR_LAB2RGB) != OT.lab2rgb_u8(c)).sum())
    assert bad == 0


def test_lab_table_header_is_current():
    assert subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_lab_tables.py"), "--check"]) == 0


def test_lab2rgb_fits_int32():
    # the kernels evaluate the 3x3 matrix in int32, as the library does; show the intermediate sums fit
    t = OT.lab_tables()
    C = np.array(t["inv"], np.int64)
    xmax = int(OT._ab_to_xz(np.array([16384 + 255 * 33])).max())      # beyond any reachable fY + a/500
    bound = (np.abs(C).reshape(3, 3) * np.array([xmax, 16384, xmax])).sum(1).max()
    assert bound < 2 ** 31


@pytest.mark.parametrize("clip,grid", [(3.0, 8), (4.0, 4), (2.0, 8), (3.0, 16), (40.0, 8), (0.0, 8)])
def test_clahe_equals_cv2(clip, grid):
    img = _tile()
    rng = np.random.default_rng(0)
    cases = [img[..., 1], img[:640, :640, 0], rng.integers(0, 256, (640, 640), dtype=np.uint8), img[:500, :701, 2],
             (rng.integers(0, 40, (333, 257)) + 100).astype(np.uint8), np.full((64, 64), 7, np.uint8)]
    for c in cases:
        c = np.ascontiguousarray(c)
        ref = cv2.createCLAHE(clipLimit=clip, tileGridSize=(grid, grid)).apply(c)
        assert np.array_equal(ref, OT.clahe_apply(c, clip, grid, grid)), (c.shape, clip, grid)


def test_clahe_variant_equals_reference_expression():
    # the literal lines of gpu_handler.py:104-110 on the reference's test tile
    img = _tile()
    for clip, grid in ((3.0, 8), (4.0, 4)):
        lab = cv2.cvtColor(img, cv2.COLOR_RGB2LAB)
        l, a, b = cv2.split(lab)
        le = cv2.createCLAHE(clipLimit=clip, tileGridSize=(grid, grid)).apply(l)
        ref = cv2.cvtColor(cv2.merge([le, a, b]), cv2.COLOR_LAB2RGB)
        assert np.array_equal(ref, OT.clahe_rgb(img, clip, grid, grid))


@pytest.mark.parametrize("factor", [2.0, 1.8, 1.4, 1.6, 1.3, 1.0, 0.5, 0.0, 0.37, 3.3])
def test_brightness_and_contrast_equal_pillow(factor):
    ramp = np.arange(256, dtype=np.uint8)[None, :, None].repeat(4, 0).repeat(3, 2)
    for arr in (ramp, _tile()[:300, :300]):
        im = Image.fromarray(arr)
        assert np.array_equal(np.array(ImageEnhance.Brightness(im).enhance(factor)), OT.brightness(arr, factor))
        assert np.array_equal(np.array(ImageEnhance.Contrast(im).enhance(factor)), OT.contrast(arr, factor))


def test_gamma_lut_equals_reference_expression():
    img = _tile()
    for gamma in (2.0, 1.5):
        ref = (np.power(img / 255.0, 1.0 / gamma) * 255.0).astype(np.uint8)      # gpu_handler.py:119-120
        assert np.array_equal(ref, OT.gamma_lut(gamma)[img])


def test_archive_chain_equals_pillow():
    img = _tile()[:256, :320]
    pil = Image.fromarray(img)
    refs = []
    s = pil
    for b in (1.4, 1.6):                                                       # gpu_handler_archive.py:80-84
        s = ImageEnhance.Brightness(s).enhance(b)
        s = ImageEnhance.Contrast(s).enhance(1.3)
        refs.append(np.array(s))
    mine = OT.archive_variations(img)
    assert np.array_equal(mine[2], refs[0]) and np.array_equal(mine[3], refs[1])
    assert len(mine) == 8


def test_process_tensors_restatement_against_torch():
    import torch
    rng = np.random.default_rng(3)
    rows = [rng.random((50, 6), dtype=np.float32) * np.float32(640) for _ in range(5)]
    for r in rows:
        r[:, 4] = rng.random(50, dtype=np.float32)
    bbox = (20.9871234, 52.2291234, 20.9880567, 52.2296891)
    got = OT.process_tensors_rows(rows, bbox, 0.3)
    # the reference's own lines (gpu_handler.py:232-253) with torch CPU tensors; torch-CPU divides, torch-CUDA multiplies
    # by the reciprocal, so the comparison is on the confidence column and to 1 ulp on the coordinates
    kept = []
    for i, r in enumerate(rows):
        b = r.copy()
        b[:, 4] *= OT.confidence_adjustment(i)
        kept.append(b[b[:, 4] > 0.3])
    comb = torch.from_numpy(np.concatenate(kept, 0))
    centers = comb[:, :2] / 640
    lons = bbox[0] + centers[:, 0] * (bbox[2] - bbox[0])
    lats = bbox[3] - centers[:, 1] * (bbox[3] - bbox[1])
    assert np.array_equal(got[:, 2], comb[:, 4].numpy())
    assert np.allclose(got[:, 0], lons.numpy(), rtol=0, atol=4e-6) and np.allclose(got[:, 1], lats.numpy(), rtol=0, atol=4e-6)


def test_oracle_views_match_golden_crcs():
    # tests/golden/tta_views.json: the reference's library calls on its own test tile (make_golden.py)
    import json
    import zlib
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "tta_views.json")))
    img = _tile()
    cur = OT.lighting_variations(img) + OT.occlusion_variations(img)
    assert [zlib.crc32(np.ascontiguousarray(v).tobytes()) for v in cur] == gold["current_views_crc32"]
    assert [zlib.crc32(np.ascontiguousarray(v).tobytes()) for v in OT.archive_variations(img)] == gold["archive_views_crc32"]


def test_blend_table_equals_pillow_for_every_constant_and_value():
    v = np.arange(256, dtype=np.uint8)[None, :]
    img = Image.fromarray(v)
    for c in range(256):
        deg = Image.new("L", (256, 1), c)
        for f in (2.0, 1.8, 1.3, 0.3, -0.5):
            assert np.array_equal(np.array(Image.blend(deg, img, f))[0], OT.blend_lut(c, f)), (c, f)


def test_product_host_tables_equal_pillow_and_numpy():
    # aerial_image_recognition_b200/tta.py builds the byte curves the device applies (b2d_tta_lut); same checks as the oracle's
    from aerial_image_recognition_b200 import tta as T
    ramp = np.arange(256, dtype=np.uint8)[None, :, None].repeat(2, 0).repeat(3, 2)
    im = Image.fromarray(ramp)
    for f in (2.0, 1.8, 1.4, 1.6, 1.3, 1.0, 0.5, 0.0, 3.3):
        assert np.array_equal(np.array(ImageEnhance.Brightness(im).enhance(f))[0, :, 0], T.brightness_lut(f)), f
        assert np.array_equal(T.brightness_lut(f), OT.blend_lut(0, f))
    img = _tile()
    for gamma in (2.0, 1.5):
        assert np.array_equal((np.power(img / 255.0, 1.0 / gamma) * 255.0).astype(np.uint8), T.gamma_lut(gamma)[img])
    # view order and weights of gpu_handler.py:94-140, :274-283 and of the archived handler
    assert [k for k, _ in T.LIGHTING_VIEWS + T.OCCLUSION_VIEWS] == ["original", "clahe", "brightness", "gamma", "clahe"]
    assert [T.confidence_adjustment(i) for i in range(6)] == [1.0, 0.95, 0.90, 0.92, 0.88, 0.85]
    assert [T.confidence_adjustment(i, 12, archive=True) for i in (0, 4, 5, 7, 8, 11, 12)] == [1.0, 1.0, 0.98, 0.98, 0.95, 0.95, 0.85]
    assert len(T.ARCHIVE_VIEWS) == 8


def test_clahe_equals_cv2_on_random_shapes_grids_and_limits():
    rng = np.random.default_rng(123)
    for it in range(80):
        h, w = int(rng.integers(17, 300)), int(rng.integers(17, 300))
        tx, ty = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 16])), int(rng.choice([1, 2, 3, 4, 5, 7, 8, 16]))
        clip = float(rng.choice([0.0, 0.5, 1.0, 2.0, 3.0, 4.0, 7.3, 40.0, 1000.0]))
        kind = it % 4
        if kind == 0:
            img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        elif kind == 1:
            img = (rng.integers(0, 30, (h, w)) + rng.integers(0, 220)).astype(np.uint8)
        elif kind == 2:
            img = np.clip(np.add.outer(np.arange(h), np.arange(w)) * 255 // (h + w) + rng.integers(-5, 5, (h, w)), 0, 255).astype(np.uint8)
        else:
            img = np.full((h, w), int(rng.integers(0, 256)), np.uint8)
        ref = cv2.createCLAHE(clipLimit=clip, tileGridSize=(tx, ty)).apply(img)
        assert np.array_equal(ref, OT.clahe_apply(img, clip, tx, ty)), (h, w, tx, ty, clip, kind)
