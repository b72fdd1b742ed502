"""Regenerates the fixtures in this directory from /root/reference (run in the build container; the GPU box has no
reference tree and only reads the committed files).

* test_tile_864.png   -- the decoded pixels of the reference's only image fixture, ``test_tile.jpg`` (864 x 864 RGB,
                         BASELINE config C1), stored losslessly so that tests do not depend on a JPEG decoder version.
* default_config.json -- ``DEFAULT_CONFIG`` imported from the reference's ``_script/config.py`` (the one reference
                         module that imports without third-party packages), tuples as lists.
* test_tile_stats.json -- size, per-channel mean / std and a CRC of the decoded tile and of its two reference resizes
                         (PIL bicubic, ``simple_detector.py:463``; cv2 linear, ``_script/gpu_handler.py:74-76``) as
                         produced by the Pillow / OpenCV of this container.
* tta_views.json     -- CRC32 of every test-time-augmentation view of the tile, made with the library calls of
                         ``_script/gpu_handler.py:94-140`` (five views) and ``_script/gpu_handler_archive.py:67-122`` (eight views):
                         cv2.cvtColor / createCLAHE, PIL ImageEnhance, np.power -- by the OpenCV / Pillow / NumPy of this container.
"""
import importlib.util
import json
import os
import zlib

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def tta_views(img, a):
    import cv2
    from PIL import ImageEnhance

    def clahe(clip, grid):                                    # gpu_handler.py:104-110
        lab = cv2.cvtColor(a, cv2.COLOR_RGB2LAB)
        l, aa, bb = cv2.split(lab)
        le = cv2.createCLAHE(clipLimit=clip, tileGridSize=(grid, grid)).apply(l)
        return cv2.cvtColor(cv2.merge([le, aa, bb]), cv2.COLOR_LAB2RGB)

    def gamma(g):                                             # gpu_handler.py:118-121
        return (np.power(a / 255.0, 1.0 / g) * 255.0).astype(np.uint8)
    current = [a, clahe(3.0, 8), np.array(ImageEnhance.Brightness(img).enhance(2.0)), gamma(2.0), clahe(4.0, 4)]
    archive = [a, np.array(ImageEnhance.Brightness(img).enhance(1.8))]
    s = img
    for b in (1.4, 1.6):                                      # gpu_handler_archive.py:79-84
        s = ImageEnhance.Brightness(s).enhance(b)
        s = ImageEnhance.Contrast(s).enhance(1.3)
        archive.append(np.array(s))
    archive += [gamma(1.5), clahe(2.0, 8), clahe(4.0, 4), clahe(3.0, 16)]
    crc = lambda v: zlib.crc32(np.ascontiguousarray(v).tobytes())
    return {"current_views_crc32": [crc(v) for v in current], "archive_views_crc32": [crc(v) for v in archive],
            "opencv": cv2.__version__, "pillow": __import__("PIL").__version__, "numpy": np.__version__}


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
    json.dump(tta_views(img, a), open(os.path.join(HERE, "tta_views.json"), "w"), indent=1)
    spec = importlib.util.spec_from_file_location("ref_config", os.path.join(REF, "_script", "config.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    json.dump(mod.DEFAULT_CONFIG, open(os.path.join(HERE, "default_config.json"), "w"), indent=1, sort_keys=True)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
