"""Regenerates the fixtures in this directory from /root/reference (run in the build container; the GPU box has no
reference tree and only reads the committed files).

* test_tile_864.png   -- the decoded pixels of the reference's only image fixture, ``test_tile.jpg`` (864 x 864 RGB,
                         BASELINE config C1), stored losslessly so that tests do not depend on a JPEG decoder version.
* default_config.json -- ``DEFAULT_CONFIG`` imported from the reference's ``_script/config.py`` (the one reference
                         module that imports without third-party packages), tuples as lists.
* test_tile_stats.json -- size, per-channel mean / std and a CRC of the decoded tile and of its two reference resizes
                         (PIL bicubic, ``simple_detector.py:463``; cv2 linear, ``_script/gpu_handler.py:74-76``) as
                         produced by the Pillow / OpenCV of this container.
"""
import importlib.util
import json
import os
import zlib

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    from PIL import Image
    import cv2
    img = Image.open(os.path.join(REF, "test_tile.jpg")).convert("RGB")
    img.save(os.path.join(HERE, "test_tile_864.png"), optimize=True)
    a = np.array(img)
    pil = np.array(img.resize((640, 640)))
    cv = cv2.resize(a, (640, 640))
    stats = {"size": list(a.shape), "mean": [round(float(v), 4) for v in a.reshape(-1, 3).mean(0)],
             "std": [round(float(v), 4) for v in a.reshape(-1, 3).std(0)], "crc32": zlib.crc32(a.tobytes()),
             "pil_bicubic_640_crc32": zlib.crc32(pil.tobytes()), "cv2_linear_640_crc32": zlib.crc32(cv.tobytes()),
             "pillow": Image.__version__ if hasattr(Image, "__version__") else "", "opencv": cv2.__version__}
    json.dump(stats, open(os.path.join(HERE, "test_tile_stats.json"), "w"), indent=1)
    spec = importlib.util.spec_from_file_location("ref_config", os.path.join(REF, "_script", "config.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    json.dump(mod.DEFAULT_CONFIG, open(os.path.join(HERE, "default_config.json"), "w"), indent=1, sort_keys=True)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
