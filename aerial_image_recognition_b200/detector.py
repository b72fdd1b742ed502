"""``CarDetector`` -- the production loop of the reference (``_script/detector.py:18-276``) around the
B200 engine: config merge, paths, tile list, batches of ``batch_size`` tiles
through ``tile_handler.fetch_batch`` -> ``gpu_handler.process_batch``, duplicate removal + checkpoint every
``interval`` tiles, final duplicate removal and ``<prefix>_results.geojson``.

Same constructor, attributes and methods as the reference.  What differs, and why:

* the tile source is ``tile_source.SyntheticTileHandler`` unless ``custom_config['tile_handler']`` supplies
  an object with ``fetch_batch`` (WMS / XYZ are HTTP clients; no network here);
* the frame's bounds come from ``custom_config['frame_bounds']`` or from the ``.shp`` header
  (``utils.shapefile_bounds``) instead of geopandas;
* errors inside ``detect`` propagate unless ``custom_config['swallow_errors']`` is true -- the reference
  prints them and returns ``None`` (``detector.py:229-231``), which hides real failures.
"""
from __future__ import annotations

import os
import time
from datetime import datetime
from typing import Dict, List, Optional

from .config import DEFAULT_CONFIG
from .gpu_handler import GPUHandler
from .tile_source import SyntheticTileHandler
from .utils import CheckpointManager, ResultsManager, TileGenerator, shapefile_bounds


class CarDetector:
    def __init__(self, base_dir, custom_config=None):
        self.base_dir = base_dir
        self.config = self._load_config(custom_config)
        self._setup_paths()
        self._initialize_components()

    # -- detector.py:36-41 ------------------------------------------------------------------------------
    def _load_config(self, custom_config=None):
        config = DEFAULT_CONFIG.copy()
        if custom_config:
            config.update(custom_config)
        return config

    # -- detector.py:43-49 ------------------------------------------------------------------------------
    def _setup_paths(self):
        frame_name = os.path.splitext(self.config['frame_path'])[0]
        self.frame_path = os.path.join(self.base_dir, 'gis', 'frames', self.config['frame_path'])
        self.output_dir = os.path.join(self.base_dir, 'output', frame_name)
        self.model_path = os.path.join(self.base_dir, 'models', self.config['model_path'])
        os.makedirs(self.output_dir, exist_ok=True)

    # -- detector.py:51-86 ------------------------------------------------------------------------------
    def _initialize_components(self):
        self.tile_handler = self.config.get('tile_handler') or SyntheticTileHandler(
            size=self.config['model_input_size'][0], seed=self.config.get('synthetic_seed', 0))
        self.gpu_handler = GPUHandler(
            model_path=self.model_path,
            confidence_threshold=self.config['confidence_threshold'],
            max_gpu_memory=self.config['max_gpu_memory'],
            output_dir=self.output_dir,
            max_batch=self.config['batch_size'],
            **self.config.get('engine_options', {}))
        self.checkpoint_manager = CheckpointManager(self.output_dir)
        self.results_manager = ResultsManager(self.output_dir, prefix=self.config['output_prefix'],
                                              duplicate_distance=self.config['duplicate_distance'], engine=self.gpu_handler.engine)

    # -- detector.py:88-115 -----------------------------------------------------------------------------
    def process_images(self, image_batch, progress_bar=None):
        return self.gpu_handler.process_batch(image_batch, queue_size=self.config['queue_size'])

    def fetch_images(self, tile_batch, progress_bar=None):
        return self.tile_handler.fetch_batch(tile_batch, progress_bar)

    # -- detector.py:117-155 ----------------------------------------------------------------------------
    def _process_batch(self, batch_tiles, processed_count, total_tiles):
        images = self.tile_handler.fetch_batch(batch_tiles, None)
        if images:
            dets = self.gpu_handler.process_batch(images, queue_size=self.config.get('queue_size', self.config['batch_size']))
            return images, dets
        return [], []

    def frame_bounds(self):
        if self.config.get('frame_bounds') is not None:
            return tuple(float(v) for v in self.config['frame_bounds'])
        return shapefile_bounds(self.frame_path)

    # -- detector.py:156-237 ----------------------------------------------------------------------------
    def detect(self, interactive=True, force_restart=True):
        try:
            tiles = TileGenerator.generate_tiles(self.frame_bounds(), self.config['tile_size_meters'], self.config['tile_overlap'])
            total_tiles = len(tiles)
            if force_restart:
                processed_count, previous = 0, []
            else:
                processed_count, previous = self.checkpoint_manager.load_checkpoint()
            all_detections = list(previous) if previous else []
            # checkpoint and duplicate removal share one interval (:183-185 hard-codes 2000 = DEFAULT_CONFIG['checkpoint_interval'])
            interval = int(self.config.get('checkpoint_interval', 2000))
            last_save = processed_count
            self.stats = {'total_tiles': total_tiles, 'start': processed_count, 'checkpoints': 0, 'seconds': 0.0}
            t0 = time.time()
            while processed_count < total_tiles:
                batch_end = min(processed_count + self.config['batch_size'], total_tiles)
                _images, batch_detections = self._process_batch(tiles[processed_count:batch_end], processed_count, total_tiles)
                if batch_detections:
                    all_detections.extend(batch_detections)
                    # The reference checkpoints `processed_count` = the START of the batch whose detections are already in
                    # `all_detections` (detector.py:209-217), so a resume re-runs that batch and appends its detections a
                    # second time.  Deliberate deviation: the cursor saved is the END of the batch.
                    if batch_end - last_save >= interval:
                        all_detections = self.results_manager.remove_duplicates(all_detections)
                        self.checkpoint_manager.save_checkpoint(processed_count=batch_end, detections=all_detections,
                                                                total_tiles=total_tiles)
                        last_save = batch_end
                        self.stats['checkpoints'] += 1
                processed_count = batch_end
            all_detections = self.results_manager.remove_duplicates(all_detections)
            out = self.results_manager.process_results(all_detections)
            self.stats['seconds'] = time.time() - t0
            return out
        except Exception as e:
            if self.config.get('swallow_errors'):
                print(f"Error in detection process: {e}")
                return None
            raise
        finally:
            if hasattr(self, 'gpu_handler'):
                self.gpu_handler.cleanup()
