"""``DEFAULT_CONFIG`` with the reference's keys and defaults (``_script/config.py:3-29``), merged by
``dict.update`` as in ``CarDetector._load_config`` (``_script/detector.py:36-41``).  Keys that configure
the network tile sources are kept so caller configs load unchanged; they are unused by the synthetic
tile source that stands in for them here (no network)."""

DEFAULT_CONFIG = {
    # tile source (WMS / XYZ in the reference; kept for config compatibility)
    'wms_url': "https://service.pdok.nl/hwh/luchtfotorgb/wms/v1_0",
    'wms_layer': 'Actueel_orthoHR',
    'wms_srs': 'EPSG:4326',
    'wms_size': (1280, 1280),
    'model_input_size': (640, 640),
    'wms_format': 'image/jpeg',
    # processing
    'tile_size_meters': 64.0,
    'confidence_threshold': 0.3,
    'tile_overlap': 0.2,
    'batch_size': 64,
    'checkpoint_interval': 2000,
    'max_gpu_memory': 2.0,
    'duplicate_distance': 0,
    'num_workers': 25,
    'queue_size': 64,
    # paths
    'frame_path': 'amsterdam.shp',
    'model_path': 'car_aerial_detection_yolo7_ITCVD_deepness.onnx',
    # output
    'output_prefix': 'detections',
}
