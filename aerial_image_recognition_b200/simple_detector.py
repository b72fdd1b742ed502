"""Drop-in for the inference half of the reference's ``simple_detector.py::SimpleDetector``
(``:26-57`` constructor attributes, ``:456-504`` detect, ``:506-538`` _process_detections,
``:540-596`` _remove_duplicates, ``:648-677`` detect_batch).

The tile-fetching half (aiohttp XYZ client, ``:59-453``) is network I/O and out of scope;
callers hand in PIL images plus the ``preview_info`` dict that ``get_image`` would have
produced (``preview_info['spatial_info']['bounds']`` = west/east/south/north,
``preview_info['image_info']['crop_size']``; ``:459-460``).

What runs where:
  PIL-bicubic resize + /255 + layout   -> K1 on device (bit-exact against Pillow)
  session.run                          -> the engine
  ``boxes[:, 4] >= 0.3``               -> fused decode + compaction kernel (row order kept)
  lon/lat from the tile bounds         -> fp64 georef kernel, reference operation order
  1 m greedy dedup in UTM              -> UTM projection + grid-hash dedup kernels
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import numpy as np
import torch

from .engine import GEO_PARAMS, Engine, geodets_to_numpy
from .gpu_handler import TileStager, _as_u8_hwc, _shape_of  # noqa: F401  (host image staging shared with GPUHandler)
from .session import InferenceSession, arch_from_model_path, resolve_weights


class SimpleDetector:
    def __init__(self, model_path, output_dir, *, arch: Optional[str] = None,
                 weights=None, max_batch: int = 8, device: int = 0, seed: int = 0, precision: str = "bf16"):
        self.zoom = 21
        self.model_size = 640
        self.confidence_threshold = 0.3
        self.output_dir = output_dir
        if output_dir:
            os.makedirs(output_dir, exist_ok=True)
        earth_circumference = 40075016.686
        self.meters_per_pixel = earth_circumference / (2 ** self.zoom) / 256
        arch = arch or arch_from_model_path(model_path)
        weights = resolve_weights(model_path, arch, weights)     # FileNotFoundError unless weights="synthetic" (session.py)
        self.engine = Engine(arch, weights=weights, max_batch=max_batch, device=device, seed=seed, imgsz=self.model_size,
                             precision=precision)
        self.model = InferenceSession(engine=self.engine)
        self._stager = TileStager(self.engine)

    # -- simple_detector.py:456-504 ----------------------------------------------------------
    def detect(self, image, preview_info):
        return self.detect_batch([image], [preview_info], batch_size=1)

    # -- simple_detector.py:648-677 ----------------------------------------------------------
    def detect_batch(self, images, preview_infos, batch_size=4):
        """The reference loops one image at a time because its ONNX graph is fixed at batch 1
        (``:649-652``); the result is the concatenation in input order, which a device batch
        reproduces."""
        eng = self.engine
        if isinstance(images, torch.Tensor):
            return self._detect_device_tiles(images, preview_infos)
        images = list(images)
        shapes = [_shape_of(im) for im in images]
        out: List[dict] = []
        i = 0
        while i < len(images):
            shape = shapes[i]
            j = i
            while j < len(images) and j - i < eng.max_batch and shapes[j] == shape:
                j += 1
            n = j - i
            tiles = self._stager.upload(images[i:j], shape)        # reused pinned staging; PIL RGB pixels without re-packing
            mode = "identity" if shape[:2] == (self.model_size, self.model_size) else "pil_bicubic"
            dets, counts = eng.infer(tiles, mode, False, self.confidence_threshold, True)
            params = np.zeros((n, GEO_PARAMS), dtype=np.float64)
            for k in range(n):
                b = preview_infos[i + k]["spatial_info"]["bounds"]
                params[k, :6] = (b["west"], b["east"], b["south"], b["north"],
                                 preview_infos[i + k]["image_info"]["crop_size"], self.model_size)
            geo = eng.georef(dets, counts, torch.from_numpy(params).to(eng.device), "bounds")
            for g in geodets_to_numpy(geo, counts):
                out.extend(self._records(g))
            i = j
        return out

    def _detect_device_tiles(self, tiles: torch.Tensor, preview_infos) -> List[dict]:
        """``images`` given as one uint8 tensor [B, H, W, 3] (device-resident tiles skip the host staging; a CPU tensor is
        copied once): same records as the list form (SURVEY.md section 8b input contract)."""
        eng = self.engine
        assert tiles.dtype == torch.uint8 and tiles.dim() == 4 and tiles.shape[3] == 3 and len(preview_infos) == tiles.shape[0]
        tiles = tiles.to(eng.device, non_blocking=True)
        mode = "identity" if tuple(tiles.shape[1:3]) == (self.model_size, self.model_size) else "pil_bicubic"
        out: List[dict] = []
        for i in range(0, tiles.shape[0], eng.max_batch):
            chunk = tiles[i:i + eng.max_batch]
            n = chunk.shape[0]
            dets, counts = eng.infer(chunk, mode, False, self.confidence_threshold, True)
            params = np.zeros((n, GEO_PARAMS), dtype=np.float64)
            for k in range(n):
                b = preview_infos[i + k]["spatial_info"]["bounds"]
                params[k, :6] = (b["west"], b["east"], b["south"], b["north"],
                                 preview_infos[i + k]["image_info"]["crop_size"], self.model_size)
            geo = eng.georef(dets, counts, torch.from_numpy(params).to(eng.device), "bounds")
            for g in geodets_to_numpy(geo, counts):
                out.extend(self._records(g))
        return out

    @staticmethod
    def _records(g) -> List[dict]:
        # whole columns -> Python floats at once (``ndarray.tolist``: the same float64 values ``float(record[field])`` gives);
        # converting record by record was 7 NumPy scalar extractions per detection, 2 of the 4 ms of a C1 call
        cols = [g[k].tolist() for k in ("x", "y", "conf", "x_img", "y_img", "x_yolo", "y_yolo")]
        return [{"lon": x, "lat": y, "confidence": c, "image": {"x": xi, "y": yi}, "yolo": {"x": xy, "y": yy}}
                for x, y, c, xi, yi, xy, yy in zip(*cols)]

    # -- simple_detector.py:506-538 ----------------------------------------------------------
    def _process_detections(self, boxes, preview_info):
        """``boxes``: host rows ``[K, >=5]`` already produced by ``model.run``; filter + georef on
        the device through the same kernels as ``detect``."""
        boxes = np.ascontiguousarray(boxes, dtype=np.float32)
        if boxes.ndim != 2 or boxes.shape[0] == 0:
            return []
        eng = self.engine
        rows = boxes
        if rows.shape[1] < 6:
            rows = np.concatenate([rows, np.zeros((rows.shape[0], 6 - rows.shape[1]), np.float32)], 1)
        rt = torch.from_numpy(rows[None]).to(eng.device)
        dets, counts = eng.postprocess(1, self.confidence_threshold, True, rows=rt)
        b = preview_info["spatial_info"]["bounds"]
        params = np.zeros((1, GEO_PARAMS), dtype=np.float64)
        params[0, :6] = (b["west"], b["east"], b["south"], b["north"], preview_info["image_info"]["crop_size"], self.model_size)
        geo = eng.georef(dets, counts, torch.from_numpy(params).to(eng.device), "bounds")
        return self._records(geodets_to_numpy(geo, counts)[0])

    # -- simple_detector.py:540-596 ----------------------------------------------------------
    def _remove_duplicates(self, detections, distance_threshold=1.0):
        if not detections:
            return []
        eng = self.engine
        lon = np.array([d["lon"] for d in detections], dtype=np.float64)
        lat = np.array([d["lat"] for d in detections], dtype=np.float64)
        # priority = position in the reference's stable descending sort (:565), handed to the kernel as its int64
        # tie-break with a constant confidence: the kernel's (conf desc, tiebreak asc) order is then exactly that sort
        ranked = sorted(range(len(detections)), key=lambda i: detections[i]["confidence"], reverse=True)
        rank = np.empty(len(detections), dtype=np.int64)
        rank[ranked] = np.arange(len(detections), dtype=np.int64)
        conf = np.zeros(len(detections), dtype=np.float32)
        utm_zone = int((detections[0]["lon"] + 180) / 6) + 1      # :546
        north = detections[0]["lat"] > 0                           # :547
        x, y = eng.utm_forward(torch.from_numpy(lon).to(eng.device), torch.from_numpy(lat).to(eng.device), utm_zone, north)
        keep = eng.dedup(x, y, torch.from_numpy(conf).to(eng.device), float(distance_threshold), inclusive=True,
                         tiebreak=torch.from_numpy(rank).to(eng.device)).cpu().numpy()
        idx = np.nonzero(keep)[0]
        # kept detections come back in descending-confidence order, input order among ties (:565, :590)
        order = sorted(idx.tolist(), key=lambda i: detections[i]["confidence"], reverse=True)
        return [detections[i] for i in order]
