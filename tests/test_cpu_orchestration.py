"""Host side of SURVEY.md section 8f-2 / 8f-3: UTM tiling (``_script/utils.py:17-65``), checkpoint and result
files (``:68-146``, ``:181-292``), the synthetic tile source.  No GPU: the kernels these classes call are
covered by ``tests/test_gpu_parity.py``."""
import json
import math
import os
import struct

import numpy as np
import pytest

from aerial_image_recognition_b200 import geo, utils as U
from aerial_image_recognition_b200.config import DEFAULT_CONFIG
from aerial_image_recognition_b200.tile_source import SyntheticTileHandler
from oracle import postproc as OP


def test_utm_forward_matches_oracle_and_inverse_round_trips():
    rng = np.random.default_rng(0)
    for zone, north in [(34, True), (31, True), (11, True), (56, False), (33, False)]:
        lon0 = (zone - 1) * 6 - 180 + 3
        lon = lon0 + rng.uniform(-3, 3, 200); lat = rng.uniform(1, 70, 200) * (1 if north else -1)
        e, n = geo.utm_forward(lon, lat, zone, north)
        eo, no = OP.utm_forward(lon, lat, zone, north)
        assert np.array_equal(e, eo) and np.array_equal(n, no)          # the same series, term for term
        lo2, la2 = geo.utm_inverse(e, n, zone, north)
        assert np.abs(lo2 - lon).max() < 1e-10 and np.abs(la2 - lat).max() < 1e-10      # < 0.02 mm
    # on the central meridian the easting is the false easting and the northing k0 * meridian arc
    e, n = geo.utm_forward(21.0, 52.2, 34, True)
    assert e == 500000.0 and abs(n - 5783283.16) < 0.01
    assert geo.utm_epsg(4.9, 52.37) == "EPSG:32631" and geo.utm_epsg(151.2, -33.86) == "EPSG:32756"


def test_generate_tiles_order_size_and_overlap():
    bounds = (20.999, 52.199, 21.003, 52.2012)           # ~270 m x 245 m near Warsaw
    tiles = U.TileGenerator.generate_tiles(bounds, 64.0, 0.2)
    zone = geo.utm_zone_of(21.001)
    (x0, x1), (y0, y1) = geo.utm_forward([bounds[0], bounds[2]], [bounds[1], bounds[3]], zone, True)
    ref = OP.generate_tiles_metric(float(x0), float(y0), float(x1), float(y1), 64.0, 0.2)      # the reference's loop in metres
    assert len(tiles) == len(ref) > 20
    for (lo0, la0, lo1, la1), (rx0, ry0, rx1, ry1) in zip(tiles, ref):
        (ex0, ex1), (ny0, ny1) = geo.utm_forward([lo0, lo1], [la0, la1], zone, True)
        assert abs(ex0 - rx0) < 1e-6 and abs(ny0 - ry0) < 1e-6 and abs(ex1 - rx1) < 1e-6 and abs(ny1 - ry1) < 1e-6
    nx = sum(1 for t in ref if t[1] == ref[0][1])        # x runs fastest
    assert ref[1][0] - ref[0][0] == pytest.approx(51.2) and ref[nx][1] - ref[0][1] == pytest.approx(51.2)
    assert ref[nx - 1][2] > float(x1)                    # the last column is not clipped (utils.py:49-54)
    assert U.TileGenerator.generate_tiles((21.0, 52.0, 21.0, 52.0), 64.0, 0.2) == []
    assert U.TileGenerator.get_utm_epsg(21.0, 52.0) == "EPSG:32634"


def test_shapefile_header_bounds(tmp_path):
    p = tmp_path / "frame.shp"
    head = struct.pack(">i", 9994) + b"\0" * 20 + struct.pack(">i", 50) + struct.pack("<ii", 1000, 5)
    head += struct.pack("<4d", 4.85, 52.33, 4.95, 52.41) + struct.pack("<4d", 0, 0, 0, 0)
    p.write_bytes(head)
    assert U.shapefile_bounds(str(p)) == (4.85, 52.33, 4.95, 52.41)
    (tmp_path / "bad.shp").write_bytes(b"\0" * 100)
    with pytest.raises(ValueError):
        U.shapefile_bounds(str(tmp_path / "bad.shp"))


def test_geojson_checkpoint_round_trip(tmp_path):
    dets = [{'lon': 21.0 + 1e-5 * i, 'lat': 52.2 - 2e-5 * i, 'confidence': 0.3 + 0.01 * i} for i in range(5)]
    cm = U.CheckpointManager(str(tmp_path))
    assert cm.load_checkpoint() == (0, [])
    cm.save_checkpoint(processed_count=128, detections=dets + [("image", "bbox")], total_tiles=400)   # stray tuples are skipped (:131-133)
    state = json.load(open(cm.state_file))
    assert set(state) == {'processed_count', 'total_tiles', 'timestamp'} and state['processed_count'] == 128 and state['total_tiles'] == 400
    fc = json.load(open(cm.data_file))
    assert fc["type"] == "FeatureCollection" and fc["crs"]["properties"]["name"].endswith("CRS84") and len(fc["features"]) == 5
    assert fc["features"][2] == {"type": "Feature", "properties": {"confidence": dets[2]['confidence']},
                                 "geometry": {"type": "Point", "coordinates": [dets[2]['lon'], dets[2]['lat']]}}
    n, back = cm.load_checkpoint()
    assert n == 128 and back == dets
    assert os.path.basename(cm.state_file) == "processing_state.json" and os.path.basename(cm.data_file) == "latest_detections.geojson"
    assert os.path.basename(U.CheckpointManager(str(tmp_path), "x").state_file) == "x_processing_state.json"
    # create_geodataframe: the three accepted shapes of utils.py:153-168
    fc = U.create_geodataframe([{'geometry': (1.0, 2.0), 'confidence': 0.5}, {'lon': 3.0, 'lat': 4.0}, {'foo': 1}, 7])
    assert [f["geometry"]["coordinates"] for f in fc["features"]] == [[1.0, 2.0], [3.0, 4.0]] and fc["features"][1]["properties"]["confidence"] == 0.0


def test_results_manager_paths_and_no_cpu_fallback(tmp_path):
    rm = U.ResultsManager(str(tmp_path / "out"), prefix="detections", duplicate_distance=2.0)
    assert rm.output_file.endswith("detections_results.geojson") and os.path.isdir(tmp_path / "out")
    assert rm.process_results([]) == [] and rm.remove_duplicates([]) == []
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rm.remove_duplicates([{'lon': 21.0, 'lat': 52.0, 'confidence': 0.5}])


def test_synthetic_tile_source_shape_and_determinism():
    h = SyntheticTileHandler(size=64)
    a = h.fetch_batch([(21.0, 52.0, 21.001, 52.0006), (21.001, 52.0, 21.002, 52.0006)])
    assert isinstance(a[0], list) and len(a[0]) == 1 and a[0][0][0].size == (64, 64) and a[0][0][1] == (21.0, 52.0, 21.001, 52.0006)
    b = h.get_single_image((21.0, 52.0, 21.001, 52.0006))
    assert np.array_equal(np.asarray(a[0][0][0]), np.asarray(b[0][0])) and not np.array_equal(np.asarray(a[0][0][0]), np.asarray(a[1][0][0]))
    assert DEFAULT_CONFIG['tile_size_meters'] == 64.0 and DEFAULT_CONFIG['tile_overlap'] == 0.2 and DEFAULT_CONFIG['batch_size'] == 64


def test_default_config_equals_the_reference_module():
    """``tests/golden/default_config.json`` is ``DEFAULT_CONFIG`` imported from the reference's own ``_script/config.py``
    (``tests/golden/make_golden.py``); the drop-in keeps every key and default."""
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "default_config.json")))
    ours = {k: (list(v) if isinstance(v, tuple) else v) for k, v in DEFAULT_CONFIG.items()}
    assert ours == ref


def test_reference_tile_fixture_and_its_reference_resizes():
    """The committed PNG holds the decoded pixels of the reference's ``test_tile.jpg``; the two resizes the reference
    applies to it (PIL bicubic, cv2 linear) are reproduced bit for bit by the host-side tables the device kernels use
    (``resample.py``; the kernels themselves are checked against these libraries in the GPU tests)."""
    import zlib
    from PIL import Image
    from aerial_image_recognition_b200 import resample as RS
    st = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "test_tile_stats.json")))
    a = np.array(Image.open(os.path.join(os.path.dirname(__file__), "golden", "test_tile_864.png")).convert("RGB"))
    assert list(a.shape) == st["size"] == [864, 864, 3] and zlib.crc32(a.tobytes()) == st["crc32"]
    assert zlib.crc32(RS.emulate_pil_bicubic(a, 640, 640).tobytes()) == st["pil_bicubic_640_crc32"]
    assert zlib.crc32(RS.emulate_cv2_linear(a, 640, 640).tobytes()) == st["cv2_linear_640_crc32"]


def test_shapefile_round_trip_and_header_bounds(tmp_path):
    import struct
    from aerial_image_recognition_b200 import utils as U
    rng = np.random.default_rng(2)
    dets = [{"lon": 21.0 + float(rng.random()) * 1e-2, "lat": 52.2 + float(rng.random()) * 1e-2, "confidence": float(np.float32(rng.random()))}
            for _ in range(37)]
    base = str(tmp_path / "cars")
    U.write_shapefile(dets, base + ".shp")
    for ext in (".shp", ".shx", ".dbf", ".prj", ".cpg"):
        assert os.path.exists(base + ext)
    assert open(base + ".cpg").read() == "UTF-8"                       # as the reference's frames (gis/frames/*.cpg)
    back = U.read_shapefile(base + ".shp")
    assert [(d["lon"], d["lat"]) for d in back] == [(d["lon"], d["lat"]) for d in dets]      # doubles, bit for bit
    assert all(abs(a["confidence"] - b["confidence"]) < 1e-15 for a, b in zip(back, dets))
    # file lengths in the headers (16-bit words), record offsets in the index, bounds in both headers
    shp, shx = open(base + ".shp", "rb").read(), open(base + ".shx", "rb").read()
    assert struct.unpack(">i", shp[24:28])[0] * 2 == len(shp) == 100 + 28 * 37
    assert struct.unpack(">i", shx[24:28])[0] * 2 == len(shx) == 100 + 8 * 37
    assert struct.unpack(">2i", shx[100 + 8 * 5:108 + 8 * 5]) == ((100 + 28 * 5) // 2, 10)
    b = U.shapefile_bounds(base + ".shp")
    assert b == (min(d["lon"] for d in dets), min(d["lat"] for d in dets), max(d["lon"] for d in dets), max(d["lat"] for d in dets))
    # the FeatureCollection form and the empty layer
    U.write_shapefile(U.create_geodataframe(dets), str(tmp_path / "fc"))
    assert U.read_shapefile(str(tmp_path / "fc")) == back
    U.write_shapefile([], str(tmp_path / "empty"))
    assert U.read_shapefile(str(tmp_path / "empty")) == []


def test_rgbx_view_of_a_pil_image_is_its_pixels_without_a_copy(tmp_path):
    """``GPUHandler.process_batch`` stages PIL RGB tiles as Pillow's own 4-byte pixels (gpu_handler._rgbx_view): the view
    must hold exactly ``np.asarray(img)`` in its first three bytes, for images built in memory and for lazily opened
    files, and must decline (None -> the np.asarray route) for every other mode."""
    from PIL import Image
    from aerial_image_recognition_b200.gpu_handler import _rgbx_view, _as_u8_hwc, _shape_of
    rng = np.random.default_rng(3)
    for h, w in ((640, 640), (864, 864), (333, 641), (1, 1)):
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        img = Image.fromarray(a)
        v = _rgbx_view(img)
        assert v is not None and v.shape == (h, w, 4) and v.dtype == np.uint8
        assert np.array_equal(v[..., :3], a) and _shape_of(img) == (h, w, 3)
    lazy = Image.open(os.path.join(os.path.dirname(__file__), "golden", "test_tile_864.png"))     # not loaded yet
    v = _rgbx_view(lazy) if lazy.mode == "RGB" else None
    if lazy.mode == "RGB":
        assert np.array_equal(v[..., :3], np.asarray(lazy))
    for mode in ("L", "RGBA", "P", "I;16", "F"):
        assert _rgbx_view(Image.new(mode, (8, 8))) is None
    assert _rgbx_view(rng.integers(0, 256, (8, 8, 3), dtype=np.uint8)) is None            # arrays are not PIL images
    big = Image.new("RGB", (5000, 4000), (1, 2, 3))                                       # 80 MB: several storage blocks
    v = _rgbx_view(big)
    assert v is None or (v.shape == (4000, 5000, 4) and (v[..., :3] == (1, 2, 3)).all())
    assert np.array_equal(_as_u8_hwc(big)[0, 0], (1, 2, 3))


def test_detection_records_from_columns_equal_the_per_record_form():
    """``SimpleDetector._records`` builds the reference's record dicts (simple_detector.py:496-502) from whole columns; the
    values must be exactly what converting record by record (``float(record[field])``) gives."""
    from aerial_image_recognition_b200.engine import GEODET_DTYPE
    from aerial_image_recognition_b200.simple_detector import SimpleDetector
    rng = np.random.default_rng(5)
    g = np.zeros(257, GEODET_DTYPE)
    for name in g.dtype.names:
        g[name] = (rng.normal(size=len(g)) * 1e3).astype(g.dtype[name])
    ref = [{"lon": float(r["x"]), "lat": float(r["y"]), "confidence": float(r["conf"]),
            "image": {"x": float(r["x_img"]), "y": float(r["y_img"])},
            "yolo": {"x": float(r["x_yolo"]), "y": float(r["y_yolo"])}} for r in g]
    got = SimpleDetector._records(g)
    assert got == ref and all(type(v) is float for v in got[0].values() if not isinstance(v, dict))
    assert SimpleDetector._records(g[:0]) == []


def test_mosaic_result_record_layout_is_what_the_device_writes():
    """``MosaicDetector._pack`` writes five 8-byte words per detection on the device and re-labels them on the host as
    ``OUT_DTYPE``; this restates the packing with CPU tensors and checks every field, including negative coordinates, the
    sign bit of the confidence's bit pattern, and the largest window id / slot of config C4."""
    import torch
    from aerial_image_recognition_b200.mosaic import MosaicDetector as MD, pack_records, unpack_records
    assert MD.OUT_DTYPE.itemsize == 8 * MD.OUT_WORDS
    n = 4096
    rng = np.random.default_rng(9)
    x = torch.from_numpy(rng.normal(size=n) * 1e7)
    y = torch.from_numpy(-np.abs(rng.normal(size=n)) * 1e7)
    conf = torch.from_numpy((rng.random(n) * 2 - 1).astype(np.float32))                 # negative values exercise bit 31
    cls = torch.from_numpy(rng.integers(0, 2, n).astype(np.int32))
    wid = torch.from_numpy(rng.integers(0, 6241, n)); wid[0] = 6240
    slot = torch.from_numpy(rng.integers(0, 300, n).astype(np.int32)); slot[0] = 299
    rec = torch.empty((n, MD.OUT_WORDS), dtype=torch.int64)
    rec[:, 0] = x.view(torch.int64); rec[:, 1] = y.view(torch.int64)
    rec[:, 2] = (conf.view(torch.int32).long() & 0xFFFFFFFF) | (cls.long() << 32)
    rec[:, 3] = wid.long(); rec[:, 4] = slot.long() & 0xFFFFFFFF
    out = rec.numpy().view(MD.OUT_DTYPE).reshape(n).copy()
    for name, col in (("x", x), ("y", y), ("conf", conf), ("cls", cls), ("window", wid), ("slot", slot)):
        assert np.array_equal(out[name], col.numpy()), name
    both = np.concatenate([out, np.zeros(0, MD.OUT_DTYPE), out[:3]])                      # what callers do with per-rank parts
    assert len(both) == n + 3 and np.array_equal(np.sort(both, order=["window", "slot"])["window"], np.sort(both["window"]))
    # the exchanged 32-byte seam record round-trips the same columns (key = window * 65536 + slot)
    key = wid * 65536 + slot.long()
    ux, uy, uc, ucls, uk = unpack_records(pack_records(x, y, conf, cls, key))
    assert torch.equal(ux, x) and torch.equal(uy, y) and torch.equal(uc, conf) and torch.equal(ucls, cls) and torch.equal(uk, key)
