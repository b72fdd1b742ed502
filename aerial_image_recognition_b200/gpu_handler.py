"""Drop-in for the reference's ``_script/gpu_handler.py::GPUHandler``.

Same constructor, attributes and methods (``gpu_handler.py:16-23``, ``:67-92``,
``:151-218``, ``:287-290``); the body is the device-memory and stream manager in
front of the C-ABI engine instead of an onnxruntime session:

* ``process_batch`` keeps the reference's input contract -- a list of *lists* whose
  element ``[0]`` is ``(PIL.Image, (lon_min, lat_min, lon_max, lat_max), _)``;
  anything else is skipped (``gpu_handler.py:156-161``) -- but runs all tiles of a
  call as device batches: cv2-linear resize when the tile is not 640x640
  (``:74-76``), ``conf >= threshold`` (``:166-170``), the ten best per tile
  (``:173``), and the ``x/640*864/864`` georeferencing (``:182-190``) in fp64;
* per-batch exceptions propagate by default.  The reference swallows them and returns
  ``[]`` (``:215-218``); pass ``swallow_errors=True`` for that behaviour.
"""
from __future__ import annotations

import gc
from typing import Dict, List, Optional

import numpy as np
import torch

from .engine import GEO_PARAMS, Engine, geodets_to_numpy
from .session import InferenceSession, arch_from_model_path, load_weights


def _as_u8_hwc(img) -> np.ndarray:
    a = np.asarray(img)
    if a.ndim == 2:
        a = np.repeat(a[..., None], 3, 2)
    if a.shape[2] == 4:
        a = a[..., :3]
    return np.ascontiguousarray(a, dtype=np.uint8)


class GPUHandler:
    def __init__(self, model_path, max_gpu_memory=5.0, confidence_threshold=0.3, output_dir=None, *,
                 arch: Optional[str] = None, weights: Optional[Dict[str, np.ndarray]] = None, max_batch: int = 64,
                 top_k: int = 10, bgr: bool = False, device: int = 0, seed: int = 0, swallow_errors: bool = False,
                 precision: str = "bf16"):
        self.model_path = model_path
        self.max_gpu_memory = max_gpu_memory
        self.confidence_threshold = confidence_threshold
        self.output_dir = output_dir
        self.top_k = top_k
        self.bgr = bgr                      # the archived code fed BGR (gpu_handler.py:145); current code feeds RGB
        self.swallow_errors = swallow_errors
        self.session = None
        self._setup_gpu()
        arch = arch or arch_from_model_path(model_path)
        if weights is None and model_path:
            weights = load_weights(model_path, arch)
        self.engine = Engine(arch, weights=weights, max_batch=max_batch, device=device, seed=seed, precision=precision)
        self.session = InferenceSession(engine=self.engine)

    def _setup_gpu(self):
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA is not available!")       # gpu_handler.py:27-28

    # -- gpu_handler.py:67-92 -------------------------------------------------------------
    def preprocess_image(self, img):
        a = _as_u8_hwc(img)
        t = torch.from_numpy(a)[None].to(self.engine.device)
        S = self.engine.imgsz
        mode = "identity" if a.shape[:2] == (S, S) else "cv2_linear"
        return self.engine.preprocess(t, mode, out="f32").cpu().numpy()

    # -- gpu_handler.py:151-218 -----------------------------------------------------------
    def process_batch(self, images, queue_size=16):
        try:
            return self._process_batch(images)
        except Exception:
            if self.swallow_errors:
                import traceback
                traceback.print_exc()
                return []
            raise

    def _process_batch(self, images) -> List[dict]:
        eng = self.engine
        items = []
        for img_set in images:
            if not img_set or not isinstance(img_set, list) or not img_set[0]:
                continue
            img, bbox, _ = img_set[0]
            items.append((_as_u8_hwc(img), tuple(float(v) for v in bbox)))
        out: List[dict] = []
        i = 0
        while i < len(items):
            # consecutive tiles of one shape form a device batch
            shape = items[i][0].shape
            j = i
            while j < len(items) and j - i < eng.max_batch and items[j][0].shape == shape:
                j += 1
            n = j - i
            host = torch.from_numpy(np.stack([it[0] for it in items[i:j]])).pin_memory()
            tiles = host.to(eng.device, non_blocking=True)
            S = eng.imgsz
            mode = "identity" if shape[:2] == (S, S) else "cv2_linear"
            dets, counts = eng.infer(tiles, mode, self.bgr, self.confidence_threshold, True, 0.0, self.top_k)
            params = np.zeros((n, GEO_PARAMS), dtype=np.float64)
            for k in range(n):
                params[k, :4] = items[i + k][1]
            geo = eng.georef(dets, counts, torch.from_numpy(params).to(eng.device), "gpuhandler")
            for g in geodets_to_numpy(geo, counts):
                for r in g:
                    out.append({"lon": float(r["x"]), "lat": float(r["y"]), "confidence": float(r["conf"])})
            i = j
        return out

    def cleanup(self):
        torch.cuda.empty_cache()
        gc.collect()
