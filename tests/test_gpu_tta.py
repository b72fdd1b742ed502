"""GPU parity tests (-m gpu) of the test-time-augmentation views, through the C ABI.

Everything here is byte / integer work, so the bar is bit-exact: against OpenCV and Pillow themselves (the libraries the
reference calls at _script/gpu_handler.py:94-140) and against the NumPy oracle (oracle/tta.py, pinned to the same
libraries on the CPU side)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from aerial_image_recognition_b200 import graph as G, synth, tta as T, weights as W
from oracle import tta as OT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def eng():
    from aerial_image_recognition_b200.engine import Engine
    g = G.build("yolov8m", imgsz=128)            # the view kernels do not depend on the network; a small plan keeps this quick
    e = Engine("yolov8m", weights=W.make_synthetic_weights(g, 0), max_batch=4, graph=g)
    yield e
    e.close()


def _tile():
    from PIL import Image
    return np.array(Image.open(os.path.join(ROOT, "tests", "golden", "test_tile_864.png")).convert("RGB"))


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_colour_conversions_equal_cv2_over_all_inputs(eng):
    import cv2
    g = np.arange(256, dtype=np.uint8)
    cube = np.zeros((256, 256, 256, 3), np.uint8)
    cube[..., 0] = g[:, None, None]
    cube[..., 1] = g[None, :, None]
    cube[..., 2] = g[None, None, :]
    flat = cube.reshape(4096, 4096, 3)
    d = _dev(flat)
    for code, cvcode in (("rgb2lab", cv2.COLOR_RGB2LAB), ("lab2rgb", cv2.COLOR_LAB2RGB)):
        got = eng.colour_convert(d, code).cpu().numpy()
        assert np.array_equal(got, cv2.cvtColor(flat, cvcode)), code
    # ragged pixel count (the byte path of the last npix mod 4 pixels)
    odd = flat[:3, :1001].copy()
    assert np.array_equal(eng.colour_convert(_dev(odd), "rgb2lab").cpu().numpy(), cv2.cvtColor(odd, cv2.COLOR_RGB2LAB))


def _clahe_cv2(img, clip, grid):
    import cv2
    lab = cv2.cvtColor(img, cv2.COLOR_RGB2LAB)                       # gpu_handler.py:104-110, literally
    l, a, b = cv2.split(lab)
    le = cv2.createCLAHE(clipLimit=clip, tileGridSize=(grid, grid)).apply(l)
    return cv2.cvtColor(cv2.merge([le, a, b]), cv2.COLOR_LAB2RGB)


@pytest.mark.parametrize("clip,grid", [(3.0, 8), (4.0, 4), (2.0, 8), (3.0, 16), (40.0, 8), (0.0, 8)])
def test_clahe_views_equal_cv2(eng, clip, grid):
    img = _tile()
    rng = np.random.default_rng(0)
    batch640 = np.stack([img[:640, :640], img[100:740, 200:840], rng.integers(0, 256, (640, 640, 3), dtype=np.uint8),
                         synth.make_tiles(1, 640, 9)[0], np.full((640, 640, 3), 77, np.uint8)])
    got = eng.tta_clahe(_dev(batch640), clip, grid).cpu().numpy()
    for k in range(len(batch640)):
        assert np.array_equal(got[k], _clahe_cv2(batch640[k], clip, grid)), k
    # other sizes: the reference's 864-px tile, and ragged sizes that take cv2's reflect-101 extension
    for a in (img, img[:500, :701].copy(), img[3:336, 5:262].copy()):
        got = eng.tta_clahe(_dev(a[None]), clip, grid).cpu().numpy()[0]
        ref = _clahe_cv2(a, clip, grid)
        assert np.array_equal(got, ref), (a.shape, int((got != ref).sum()))
        assert np.array_equal(got, OT.clahe_rgb(a, clip, grid, grid))


def test_clahe_views_equal_cv2_on_random_shapes_grids_and_limits(eng):
    import cv2
    rng = np.random.default_rng(321)
    for it in range(48):
        h, w = int(rng.integers(17, 300)), int(rng.integers(17, 300))
        if it % 3 == 0:
            h, w = 4 * (h // 4 + 1), 8 * (w // 8 + 1)           # shapes that take the aligned 4-pixel histogram path
        tx, ty = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 16])), int(rng.choice([1, 2, 3, 4, 5, 7, 8, 16]))
        clip = float(rng.choice([0.0, 0.5, 1.0, 2.0, 3.0, 4.0, 7.3, 40.0, 1000.0]))
        n = 1 + it % 3
        if it % 4 == 3:
            a = (rng.integers(0, 30, (n, h, w, 3)) + rng.integers(0, 220)).astype(np.uint8)
        else:
            a = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
        got = eng.tta_clahe(_dev(a), clip, (tx, ty)).cpu().numpy()
        for k in range(n):
            lab = cv2.cvtColor(a[k], cv2.COLOR_RGB2LAB)
            l, aa, bb = cv2.split(lab)
            le = cv2.createCLAHE(clipLimit=clip, tileGridSize=(tx, ty)).apply(l)
            ref = cv2.cvtColor(cv2.merge([le, aa, bb]), cv2.COLOR_LAB2RGB)
            assert np.array_equal(got[k], ref), (h, w, tx, ty, clip, n, k, int((got[k] != ref).sum()))


def test_brightness_gamma_contrast_equal_pillow_and_numpy(eng):
    from PIL import Image, ImageEnhance
    img = _tile()
    for a in (img[:640, :640].copy(), img[:333, :257].copy()):        # 16-byte vector path, byte path
        d = _dev(a[None])
        pil = Image.fromarray(a)
        for f in (2.0, 1.8, 1.4, 1.6, 0.5, 3.3):
            assert np.array_equal(eng.tta_lut(d, T.brightness_lut(f)).cpu().numpy()[0], np.array(ImageEnhance.Brightness(pil).enhance(f)))
            assert np.array_equal(eng.tta_contrast(d, f).cpu().numpy()[0], np.array(ImageEnhance.Contrast(pil).enhance(f))), f
        for gamma in (2.0, 1.5):
            ref = (np.power(a / 255.0, 1.0 / gamma) * 255.0).astype(np.uint8)      # gpu_handler.py:119-120
            assert np.array_equal(eng.tta_lut(d, T.gamma_lut(gamma)).cpu().numpy()[0], ref)
    # per-image means in one batched call
    batch = np.stack([img[:256, :256], 255 - img[:256, :256], img[300:556, 300:556] // 3])
    got = eng.tta_contrast(_dev(batch), 1.3).cpu().numpy()
    for k in range(3):
        assert np.array_equal(got[k], np.array(ImageEnhance.Contrast(Image.fromarray(batch[k])).enhance(1.3)))
        assert np.array_equal(got[k], OT.contrast(batch[k], 1.3))


def test_view_sets_equal_oracle_and_libraries(eng):
    tiles = synth.make_tiles(3, 640, 21)
    d = _dev(tiles)
    cur = eng.tta_views(d, T.LIGHTING_VIEWS + T.OCCLUSION_VIEWS)
    arc = eng.tta_views(d, T.ARCHIVE_VIEWS)
    assert len(cur) == 5 and len(arc) == 8
    for k in range(3):
        ref = OT.lighting_variations(tiles[k]) + OT.occlusion_variations(tiles[k])
        for i in range(5):
            assert np.array_equal(cur[i][k].cpu().numpy(), ref[i]), (k, i)
        ref = OT.archive_variations(tiles[k])
        for i in range(8):
            assert np.array_equal(arc[i][k].cpu().numpy(), ref[i]), (k, i)


def test_views_of_the_reference_tile_match_golden_crcs(eng):
    # tests/golden/tta_views.json was made with the reference's library calls on test_tile.jpg (864 x 864, no resize)
    import json
    import zlib
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "tta_views.json")))
    d = _dev(_tile()[None])
    crc = lambda t: zlib.crc32(t[0].cpu().numpy().tobytes())
    assert [crc(v) for v in eng.tta_views(d, T.LIGHTING_VIEWS + T.OCCLUSION_VIEWS)] == gold["current_views_crc32"]
    assert [crc(v) for v in eng.tta_views(d, T.ARCHIVE_VIEWS)] == gold["archive_views_crc32"]


def test_conf_scale_strict_filter_and_float32_georef(eng):
    rng = np.random.default_rng(11)
    A = 500
    rows = [(rng.random((A, 6), dtype=np.float32) * np.float32(640)) for _ in range(5)]
    for r in rows:
        r[:, 4] = rng.random(A, dtype=np.float32)
        r[:7, 4] = np.float32(0.3) / np.float32(0.95)            # lands on / next to the threshold after scaling
    bbox = (20.9871234, 52.2291234, 20.9880567, 52.2296891)
    params = np.zeros((1, 16)); params[0, :4] = bbox
    got = []
    for i, r in enumerate(rows):
        eng.set_conf_scale(T.confidence_adjustment(i))
        dets, counts = eng.postprocess(1, 0.3, False, rows=_dev(r[None]))
        eng.set_conf_scale(1.0)
        geo = eng.georef(dets, counts, _dev(params), "tensor_f32")
        from aerial_image_recognition_b200.engine import geodets_to_numpy
        g = geodets_to_numpy(geo, counts)[0]
        got.append(np.stack([g["x"].astype(np.float32), g["y"].astype(np.float32), g["conf"]], 1))
    got = np.concatenate(got, 0)
    assert np.array_equal(got, OT.process_tensors_rows(rows, bbox, 0.3))
    # the reference's own lines (gpu_handler.py:232-253) on CUDA tensors
    kept = []
    for i, r in enumerate(rows):
        b = r.copy()
        b[:, 4] *= T.confidence_adjustment(i)
        kept.append(b[b[:, 4] > 0.3])
    bt = torch.from_numpy(np.concatenate(kept, 0)).cuda()
    centers = bt[:, :2] / 640
    lons = bbox[0] + (centers[:, 0] * (bbox[2] - bbox[0]))
    lats = bbox[3] - (centers[:, 1] * (bbox[3] - bbox[1]))
    ref = torch.stack([lons, lats, bt[:, 4]], dim=1).cpu().numpy()
    assert ref.dtype == np.float32 and np.array_equal(got, ref)


def test_gpu_handler_tta_paths_agree_with_reference_restatement():
    from aerial_image_recognition_b200.gpu_handler import GPUHandler
    g = G.build("yolov7")
    w = W.make_synthetic_weights(g, 0)
    h = GPUHandler("car_aerial_detection_yolo7_ITCVD_deepness.onnx", confidence_threshold=0.3, weights=w, max_batch=4)
    tiles = synth.make_tiles(3, 640, 33)
    bboxes = [(21.0 + 0.001 * i, 52.0, 21.0006 + 0.001 * i, 52.0004) for i in range(3)]
    # reference flow: views as float32 BGR tensors -> _process_tensors
    tensor_batch = [(h.preprocess_variations(tiles[k]), bboxes[k]) for k in range(3)]
    assert len(tensor_batch[0][0]) == 5 and tensor_batch[0][0][0].shape == (3, 640, 640)
    # _prepare_tensor is the reference's expression on the device (gpu_handler.py:142-149)
    ref0 = (torch.from_numpy(np.ascontiguousarray(tiles[0][..., ::-1])).cuda().to(torch.float32).permute(2, 0, 1) / 255.0)
    assert torch.equal(tensor_batch[0][0][0], ref0)
    a = h._process_tensors(tensor_batch)
    b = h.process_batch_tta([[(tiles[k], bboxes[k], None)] for k in range(3)])
    assert a == b and len(a) > 0
    # restatement: rows of every view from the session (the network itself is covered by test_gpu_parity), then the
    # oracle's filter / scale / float32 georeferencing
    ref = []
    name = h.session.get_inputs()[0].name
    for k in range(3):
        views = OT.lighting_variations(tiles[k]) + OT.occlusion_variations(tiles[k])
        rows = []
        for v in views:
            x = np.expand_dims((v[..., ::-1].astype(np.float32) / 255.0).transpose(2, 0, 1), 0)
            rows.append(h.session.run(None, {name: np.ascontiguousarray(x)})[0][0])
        for lon, lat, conf in OT.process_tensors_rows(rows, bboxes[k], 0.3):
            ref.append({"lon": float(lon), "lat": float(lat), "confidence": float(conf)})
    assert a == ref
    h.cleanup()
    h.engine.close()


def test_views_at_full_batch_are_per_image(eng):
    # BASELINE size (64 x 640^2): every image of a batched call equals the same image processed alone
    tiles = synth.make_tiles(64, 640, 77)
    d = _dev(tiles)
    views = eng.tta_views(d, T.LIGHTING_VIEWS + T.OCCLUSION_VIEWS)
    for k in (0, 17, 63):
        single = eng.tta_views(d[k:k + 1], T.LIGHTING_VIEWS + T.OCCLUSION_VIEWS)
        for i in range(5):
            assert torch.equal(views[i][k], single[i][0]), (k, i)
    ident = np.arange(256, dtype=np.uint8)
    assert torch.equal(eng.tta_lut(d, ident), d)
    assert torch.equal(eng.tta_contrast(d, 1.0), d)
