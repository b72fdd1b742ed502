"""CPU tests (-m "not gpu"): the oracle against hand-computed cases and library oracles
(Pillow, OpenCV, torchvision), the graph wiring, the host tables, and the C-ABI surface."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest
import torch

from aerial_image_recognition_b200 import graph as G, resample as R, synth, weights as W
from oracle import postproc as OP
from oracle.yolo_torch import make_oracle
from _ir_cpu import run_graph_cpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- graph / weights ---------------------------------------------------------------------
def test_v8m_matches_training_log_counts():
    # x_arch/01_train_tokyo.ipynb:1 (cell 15 output): fused 23,203,990 params, 67.4 GFLOPs
    g = G.build("yolov8m")
    assert g.fused_param_count() == 23_203_990
    assert abs(2 * g.macs_per_tile() / 1e9 - 67.43) < 0.01
    assert len(g.wshapes) == 89 and g.head["anchors_total"] == 8400


def test_v7_canonical_counts():
    g = G.build("yolov7")
    assert abs(g.macs_per_tile() / 1e9 - 51.576) < 0.001
    assert len(g.wshapes) == 92 and g.head["anchors_total"] == 25200


@pytest.mark.parametrize("arch", ["yolov8m", "yolov7"])
def test_offset_write_graph_equals_conventional_oracle(arch):
    g = G.build(arch, imgsz=128)
    w = W.make_synthetic_weights(g, 1)
    x = torch.from_numpy(synth.make_tiles(2, 128, 3).astype(np.float32) / 255).permute(0, 3, 1, 2)
    for emu in (False, True):
        raw = make_oracle(arch, w, emulate_bf16=emu).raw_head(x)
        bufs = run_graph_cpu(g, w, x, emulate_bf16=emu)
        for i, lv in enumerate(g.head["levels"]):
            got = bufs[lv["buf"]][:, :raw[i].shape[1]]
            # same arithmetic, but sliced (non-contiguous) inputs may take another oneDNN path;
            # with bf16 rounding a last-bit difference can flip a rounding and propagate
            err = (got - raw[i]).abs().max().item() / raw[i].abs().max().item()
            assert err < (5e-2 if emu else 1e-4), (arch, emu, i, err)


def test_xunet_stand_in_graph_equals_conventional_oracle():
    """Config C5's declared stand-in (graph.build_xunet): the offset-write op list (skips written straight into the decoder's
    concat buffers, stride 2 as depthwise + max-pool) against the conventionally written NCHW module of oracle/xunet_torch.py."""
    from oracle.xunet_torch import XUnetOracle
    g = G.build("xunet", imgsz=64)
    assert g.head["kind"] == "seg" and g.bufs["logits"].f32 and sum(op.kind == "dwconv" for op in g.ops) == 16
    full = G.build("xunet")
    assert full.imgsz == 256 and full.macs_per_tile() == 3124379648 and full.fused_param_count() == 5472372
    w = W.make_synthetic_weights(g, 1)
    x = torch.from_numpy(synth.make_tiles(2, 64, 3).astype(np.float32) / 255).permute(0, 3, 1, 2)
    for emu in (False, True):
        ref = XUnetOracle(w, emulate_bf16=emu).logits(x)
        got = run_graph_cpu(g, w, x, emulate_bf16=emu)["logits"][:, :4]
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        assert err < (5e-2 if emu else 1e-4), (emu, err)
    labels, conf = XUnetOracle(w).forward(x)
    assert labels.shape == (2, 64, 64) and labels.dtype == torch.uint8 and int(labels.max()) <= 3
    assert float(conf.min()) >= 0.25 - 1e-6 and float(conf.max()) <= 1.0 + 1e-6          # softmax maximum of four classes


def test_weights_are_deterministic_and_bf16_representable():
    g = G.build("yolov8m", imgsz=64)
    a = W.make_synthetic_weights(g, 0)
    b = W.make_synthetic_weights(g, 0)
    assert W.weights_fingerprint(a) == W.weights_fingerprint(b)
    for k, v in a.items():
        if k.endswith(".weight"):
            assert np.array_equal(W.round_to_bf16(v), v)
    assert W.weights_fingerprint(W.make_synthetic_weights(g, 1)) != W.weights_fingerprint(a)


# ---- resize: emulation of the kernels' integer arithmetic vs the libraries -------------------
@pytest.mark.parametrize("shape", [(864, 864), (1280, 1280), (1000, 1300), (777, 900), (640, 864)])
def test_resize_emulation_bit_exact(shape):
    import cv2
    from PIL import Image
    rng = np.random.default_rng(shape[0])
    a = rng.integers(0, 256, (*shape, 3), dtype=np.uint8)
    assert np.array_equal(np.array(Image.fromarray(a).resize((640, 640))), R.emulate_pil_bicubic(a, 640, 640))
    assert np.array_equal(cv2.resize(a, (640, 640)), R.emulate_cv2_linear(a, 640, 640))


def test_resize_on_reference_tile():
    import cv2
    from PIL import Image
    a = np.array(Image.open(os.path.join(ROOT, "tests/golden/test_tile_864.png")).convert("RGB"))
    assert a.shape == (864, 864, 3)
    assert np.array_equal(np.array(Image.fromarray(a).resize((640, 640))), R.emulate_pil_bicubic(a, 640, 640))
    assert np.array_equal(cv2.resize(a, (640, 640)), R.emulate_cv2_linear(a, 640, 640))


def test_letterbox_geometry():
    assert R.letterbox_geometry(1200, 1200)[:4] == (640, 640, 0, 0)
    nw, nh, left, top, r = R.letterbox_geometry(300, 1200)
    assert (nw, nh, left, top) == (640, 160, 0, 240)


# ---- C ABI surface (no GPU needed) -----------------------------------------------------------
def test_library_exports_every_declared_symbol(lib):
    from aerial_image_recognition_b200 import _lib
    hdr = open(os.path.join(ROOT, "include/b2det.h")).read()
    declared = set(re.findall(r"\b(b2d_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in b2det.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.b2d_version() == 101


def test_create_fails_loudly_without_gpu(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.b2d_create(0, 4, C.byref(h)) < 0
    assert b"no CPU fallback" in lib.b2d_last_error()
    from aerial_image_recognition_b200.engine import Engine
    from aerial_image_recognition_b200._lib import B2DError
    with pytest.raises(B2DError):
        Engine("yolov8m", max_batch=1)


def test_entry_points_reject_bad_arguments_with_a_message(lib):
    # status < 0 and a text in b2d_last_error(), never a crash: no engine handle is needed to see the argument checks
    null = C.c_void_p(0)
    calls = [
        ("b2d_tta_clahe", (null, null, 1, 640, 640, 3.0, 8, 8, null, null), b"tta_clahe"),
        ("b2d_tta_lut", (null, null, 1, 100, null, 0, null, null), b"tta_lut"),
        ("b2d_tta_contrast", (null, null, 1, 640, 640, 1.3, null, null), b"tta_contrast"),
        ("b2d_colour_convert", (null, null, 16, 0, null, null), b"colour_convert"),
        ("b2d_set_conf_scale", (null, 0.95), b"set_conf_scale"),
        ("b2d_detect_host", (null, null, 1, 640, 640, 0, 0, 0.3, 1, 0.0, 0, 300, 0, null, null, null, 300, null), b"detect_host"),
        ("b2d_preprocess", (null, null, 1, 640, 640, 1920, 1228800, 0, 0, 0, null, null), b"preprocess"),
        ("b2d_georef", (null, null, null, 1, 1, 0, null, null, null), b"georef"),
    ]
    for name, args, word in calls:
        assert getattr(lib, name)(*args) < 0, name
        assert word in lib.b2d_last_error(), (name, lib.b2d_last_error())


@pytest.mark.parametrize("mode,fn", [(2, "pil"), (1, "cv2")])
@pytest.mark.parametrize("sizes", [(864, 640), (1280, 640), (1300, 640), (300, 640), (640, 640)])
def test_c_resize_tables_equal_python_tables(lib, mode, fn, sizes):
    i, o = sizes
    k = C.c_int()
    lib.b2d_resize_table(mode, i, o, None, None, C.byref(k))
    b = np.zeros((o, 2), np.int32); c = np.zeros((o, k.value), np.int32)
    assert lib.b2d_resize_table(mode, i, o, b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), C.byref(k)) == 0
    if fn == "pil":
        pb, pk, ks = R.pil_bicubic_table(i, o)
        assert ks == k.value and np.array_equal(b, pb) and np.array_equal(c, pk)
    else:
        co, cc = R.cv2_linear_table(i, o)
        assert np.array_equal(b, co) and np.array_equal(c, cc.astype(np.int32))


# ---- post-processing oracle ---------------------------------------------------------------------
def test_filter_is_inclusive_on_column_4():
    rows = np.array([[1, 2, 3, 4, 0.3, 0.9], [1, 2, 3, 4, np.nextafter(np.float32(0.3), np.float32(0)), 0.99]], np.float32)
    assert len(OP.filter_rows(rows, 0.3)) == 1 and len(OP.filter_rows(rows, 0.3, inclusive=False)) == 0


def test_nms_loop_matches_torchvision():
    import torchvision
    rng = np.random.default_rng(0)
    for _ in range(5):
        n = 400
        c = rng.uniform(0, 300, (n, 2)); wh = rng.uniform(10, 60, (n, 2))
        b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
        s = rng.random(n).astype(np.float32); s[:20] = 0.5
        ref = torchvision.ops.nms(torch.from_numpy(b), torch.from_numpy(s), 0.5).numpy()
        assert np.array_equal(ref, OP.nms_greedy_reference(b, s, 0.5))


def test_nms_iou_exactly_at_threshold_is_kept():
    b = np.array([[0, 0, 10, 10], [0, 0, 10, 5]], np.float32)    # IoU = 0.5 exactly
    assert len(OP.nms_greedy_reference(b, np.array([0.9, 0.8], np.float32), 0.5)) == 2


def test_georef_is_float64_not_numpy2_float32():
    # SURVEY.md Appendix B.5 / D.6
    lon, lat, xi, yi = OP.georef_bounds(np.float32(123.456), np.float32(500.25), 21.0, 21.065, 52.0, 52.04)
    assert lon == 21.0 + (float(np.float32(123.456)) / 640) * (21.065 - 21.0)
    # the same source evaluated with NumPy >= 2 scalar promotion stays in float32 and differs
    xf32 = np.float32(123.456) / np.float32(640)
    lon32 = float(np.float32(21.0) + xf32 * np.float32(21.065 - 21.0))
    assert lon32 != lon and abs(lon32 - lon) < 1e-5
    assert isinstance(lon, float)
    l2, _ = OP.georef_gpuhandler(np.float32(123.456), np.float32(500.25), 21.0, 52.0, 21.065, 52.04)
    assert abs(l2 - lon) < 1e-12


def test_georef_affine_sample_transform():
    gt = (2335637.62, 0.21, 0.0, 6845688.78, 0.0, -0.21)       # x_arch/02_analyze_images:1 (cell 3 output)
    assert OP.georef_affine(100, 200, gt) == (2335637.62 + 100 * 0.21 + 200 * 0.0, 6845688.78 + 100 * 0.0 + 200 * -0.21)


def test_dedup_grid_equals_bruteforce_and_is_order_stable():
    rng = np.random.default_rng(2)
    for incl in (True, False):
        x = rng.uniform(0, 30, 600); y = rng.uniform(0, 30, 600)
        conf = rng.random(600).astype(np.float32); conf[:50] = 0.5
        a = OP.dedup_greedy(x, y, conf, 1.0, incl); b = OP.dedup_bruteforce(x, y, conf, 1.0, incl)
        assert np.array_equal(a, b)
    # exact distance == thr: inclusive removes, strict keeps
    x = np.array([0.0, 1.0]); y = np.array([0.0, 0.0]); c = np.array([0.9, 0.8], np.float32)
    assert len(OP.dedup_greedy(x, y, c, 1.0, True)) == 1 and len(OP.dedup_greedy(x, y, c, 1.0, False)) == 2
    assert len(OP.dedup_greedy(np.zeros(0), np.zeros(0), np.zeros(0, np.float32), 1.0)) == 0


def test_utm_forward_known_point():
    # CN Tower: 43.642566 N, 79.387139 W -> UTM 17T 630084 E 4833438 N (published example)
    z, north = OP.utm_zone(-79.387139, 43.642566)
    assert (z, north) == (17, True)
    e, n = OP.utm_forward(-79.387139, 43.642566, z, north)
    assert abs(float(e) - 630084) < 1.0 and abs(float(n) - 4833438) < 1.0


def test_sliding_window_grid_of_the_40k_mosaic():
    w = OP.sliding_windows(40000, 40000, 640, 512)
    assert len(w) == 79 * 79 == 6241
    assert w[0] == (0, 0, 640, 640) and w[1] == (512, 0, 1152, 640)          # x inner
    assert w[78] == (39936, 0, 40000, 640) and w[79] == (0, 512, 640, 1152)  # clipped last column


def test_generate_tiles_metric_accumulates_by_addition():
    t = OP.generate_tiles_metric(0.0, 0.0, 130.0, 60.0, 64.0, 0.2)
    step = 64.0 * (1 - 0.2)
    assert t[0] == (0.0, 0.0, 64.0, 64.0) and t[1][0] == step and t[2][0] == step + step
    assert len(t) == 3 * 2


def test_utm_forward_against_published_meridian_arcs():
    # On the central meridian the northing is k0 x the meridian arc.  WGS84 arcs from the geodesy literature: quarter meridian
    # 10 001 965.729 m, equator -> 45 deg 4 984 944.378 m, equator -> 30 deg 3 320 113.398 m.  Pins the Krueger series of the
    # oracle and of the product's host projection (geo.py) where pyproj itself cannot be run.
    from aerial_image_recognition_b200 import geo
    k0 = 0.9996
    for fwd in (OP.utm_forward, geo.utm_forward):
        e, n = fwd(3.0, 0.0, 31, True)
        assert float(e) == 500000.0 and float(n) == 0.0
        e, n = fwd(3.0, 45.0, 31, True)
        assert abs(float(e) - 500000.0) < 1e-6 and abs(float(n) - k0 * 4984944.378) < 1e-3
        e, n = fwd(3.0, 89.999999, 31, True)
        assert abs(float(n) - k0 * 10001965.729) < 0.2
        e, n = fwd(9.0, -30.0, 32, False)
        assert abs(float(n) - (10000000.0 - k0 * 3320113.398)) < 1e-3
