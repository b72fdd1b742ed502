"""Tiling, checkpoints and result files of the orchestration layer:
``TileGenerator``, ``CheckpointManager``, ``ResultsManager``, ``create_geodataframe`` with the names,
arguments and on-disk formats of ``_script/utils.py:15-292``.

geopandas / shapely / pyproj are not installable here, so

* a "GeoDataFrame" is a plain ``FeatureCollection`` dict (``Point`` features with a ``confidence``
  property, CRS84) written and read as GeoJSON -- the layout GDAL's GeoJSON driver produces for
  ``gdf.to_file(..., driver='GeoJSON')``; files are compared as parsed JSON, not byte for byte;
* the projection is ``geo.py`` (Krueger series) on the host for tile corners and the device kernels
  (``b2d_utm_forward`` + ``b2d_dedup``) for the detections;
* a frame's ``total_bounds`` is read from the 100-byte header of the ``.shp`` file.
"""
from __future__ import annotations

import json
import os
import struct
from datetime import datetime
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import geo

__all__ = ['TileGenerator', 'CheckpointManager', 'ResultsManager', 'create_geodataframe', 'write_geojson', 'read_geojson',
           'shapefile_bounds', 'write_shapefile', 'read_shapefile']

CRS84 = {"type": "name", "properties": {"name": "urn:ogc:def:crs:OGC:1.3:CRS84"}}


# ---- frame bounds ------------------------------------------------------------------------------------------
def shapefile_bounds(path: str) -> Tuple[float, float, float, float]:
    """``gpd.read_file(frame).total_bounds`` (``_script/detector.py:163``) = the bounding box every ESRI
    shapefile carries in its main-file header: bytes 36..68, little-endian doubles Xmin, Ymin, Xmax, Ymax
    (ESRI Shapefile Technical Description, 1998, table 1)."""
    with open(path, "rb") as f:
        head = f.read(100)
    if len(head) < 100 or struct.unpack(">i", head[:4])[0] != 9994:
        raise ValueError(f"{path}: not an ESRI shapefile (bad file code)")
    return struct.unpack("<4d", head[36:68])


# ---- ESRI shapefile output -----------------------------------------------------------------------------
# The reference's frames and earlier results are shapefiles with .cpg "UTF-8" sidecars (gis/frames/*.cpg,
# gis/shp/x_arch/centroidy/); ``gdf.to_file(path)`` without a driver writes this format.  Point layer, one
# ``confidence`` field, WGS84 -- the file set GDAL's "ESRI Shapefile" driver produces for such a frame: .shp / .shx
# (ESRI Shapefile Technical Description, 1998), .dbf (dBASE III, Real fields as N(24,15)), .prj, .cpg.
WGS84_PRJ = ('GEOGCS["GCS_WGS_1984",DATUM["D_WGS_1984",SPHEROID["WGS_1984",6378137.0,298.257223563]],'
             'PRIMEM["Greenwich",0.0],UNIT["Degree",0.0174532925199433]]')


def _shp_header(file_words: int, bbox) -> bytes:
    return (struct.pack(">7i", 9994, 0, 0, 0, 0, 0, file_words) + struct.pack("<2i", 1000, 1)
            + struct.pack("<8d", bbox[0], bbox[1], bbox[2], bbox[3], 0.0, 0.0, 0.0, 0.0))


def write_shapefile(detections, path: str) -> None:
    """Point shapefile of ``[{'lon', 'lat', 'confidence'}, ...]`` (or of a FeatureCollection from
    ``create_geodataframe``).  ``path`` may end in ``.shp``; the four sidecars are written next to it."""
    if isinstance(detections, dict):
        detections = [{"lon": f["geometry"]["coordinates"][0], "lat": f["geometry"]["coordinates"][1],
                       "confidence": f["properties"]["confidence"]} for f in detections.get("features", [])]
    base = path[:-4] if path.lower().endswith(".shp") else path
    pts = [(float(d["lon"]), float(d["lat"]), float(d["confidence"])) for d in detections]
    n = len(pts)
    bbox = (min(p[0] for p in pts), min(p[1] for p in pts), max(p[0] for p in pts), max(p[1] for p in pts)) if n else (0.0,) * 4
    with open(base + ".shp", "wb") as shp, open(base + ".shx", "wb") as shx:
        shp.write(_shp_header(50 + 14 * n, bbox))
        shx.write(_shp_header(50 + 4 * n, bbox))
        for i, (x, y, _) in enumerate(pts):
            shx.write(struct.pack(">2i", 50 + 14 * i, 10))                     # offset and content length in 16-bit words
            shp.write(struct.pack(">2i", i + 1, 10) + struct.pack("<i2d", 1, x, y))
    now = datetime.now()
    with open(base + ".dbf", "wb") as dbf:
        dbf.write(struct.pack("<4BIHH20x", 3, now.year - 1900, now.month, now.day, n, 32 + 32 + 1, 1 + 24))
        dbf.write(b"confidence".ljust(11, b"\0") + b"N" + b"\0" * 4 + bytes([24, 15]) + b"\0" * 14)
        dbf.write(b"\r")
        for _, _, c in pts:
            dbf.write(b" " + f"{c:24.15f}".encode("ascii")[-24:])
        dbf.write(b"\x1a")
    with open(base + ".prj", "w") as f:
        f.write(WGS84_PRJ)
    with open(base + ".cpg", "w") as f:
        f.write("UTF-8")


def read_shapefile(path: str) -> List[Dict]:
    """Reads a Point shapefile with a ``confidence`` field back into detection dicts (tests, resume)."""
    base = path[:-4] if path.lower().endswith(".shp") else path
    with open(base + ".shp", "rb") as f:
        data = f.read()
    if struct.unpack(">i", data[:4])[0] != 9994 or struct.unpack("<i", data[32:36])[0] != 1:
        raise ValueError(f"{path}: not a Point shapefile")
    pts, off = [], 100
    while off < len(data):
        _, words = struct.unpack(">2i", data[off:off + 8])
        stype, x, y = struct.unpack("<i2d", data[off + 8:off + 28])
        if stype == 1:
            pts.append((x, y))
        off += 8 + 2 * words
    with open(base + ".dbf", "rb") as f:
        d = f.read()
    nrec, hlen, rlen = struct.unpack("<IHH", d[4:12])
    fields, o, pos = [], 32, 1
    while d[o] != 0x0D:
        name = d[o:o + 11].split(b"\0")[0].decode()
        fields.append((name, pos, d[o + 16]))
        pos += d[o + 16]
        o += 32
    fpos = {name: (p, ln) for name, p, ln in fields}["confidence"]
    conf = [float(d[hlen + i * rlen + fpos[0]: hlen + i * rlen + fpos[0] + fpos[1]]) for i in range(nrec)]
    return [{"lon": x, "lat": y, "confidence": c} for (x, y), c in zip(pts, conf)]


# ---- _script/utils.py:15-65 --------------------------------------------------------------------------------
class TileGenerator:
    @staticmethod
    def get_utm_epsg(lon, lat):
        return geo.utm_epsg(lon, lat)

    @staticmethod
    def generate_tiles(bounds, tile_size_meters, overlap=0.1):
        """Square tiles in the UTM zone of the bounds' centre, returned as WGS84 boxes
        ``(lon0, lat0, lon1, lat1)``: y outer, x inner, unclipped, the step accumulated by repeated addition
        (``_script/utils.py:43-63``) -- that order is the tile index of the whole run (checkpoints count in it)."""
        minx, miny, maxx, maxy = (float(v) for v in bounds)
        center_lon, center_lat = (minx + maxx) / 2, (miny + maxy) / 2
        zone, north = geo.utm_zone_of(center_lon), not (center_lat < 0)
        (ux0, ux1), (uy0, uy1) = geo.utm_forward([minx, maxx], [miny, maxy], zone, north)
        ux0, ux1, uy0, uy1 = float(ux0), float(ux1), float(uy0), float(uy1)
        xs, ys = [], []
        y = uy0
        while y < uy1:
            x = ux0
            while x < ux1:
                xs.append(x); ys.append(y)
                x += tile_size_meters * (1 - overlap)
            y += tile_size_meters * (1 - overlap)
        if not xs:
            return []
        xs = np.asarray(xs); ys = np.asarray(ys)
        lon0, lat0 = geo.utm_inverse(xs, ys, zone, north)
        lon1, lat1 = geo.utm_inverse(xs + tile_size_meters, ys + tile_size_meters, zone, north)
        return [(float(a), float(b), float(c), float(d)) for a, b, c, d in zip(lon0, lat0, lon1, lat1)]


# ---- GeoJSON in place of GeoDataFrame ----------------------------------------------------------------------
def create_geodataframe(detections) -> Dict:
    """``create_geodataframe`` (``_script/utils.py:148-179``): dict detections -> point features with a
    ``confidence`` property; entries that are not detection dicts are skipped, a missing confidence is 0.0."""
    feats = []
    for d in detections:
        if not isinstance(d, dict):
            continue
        if 'geometry' in d and 'confidence' in d:
            x, y = d['geometry']
            conf = d['confidence']
        elif 'lon' in d and 'lat' in d:
            x, y, conf = d['lon'], d['lat'], d.get('confidence', 0.0)
        else:
            continue
        feats.append({"type": "Feature", "properties": {"confidence": float(conf)},
                      "geometry": {"type": "Point", "coordinates": [float(x), float(y)]}})
    return {"type": "FeatureCollection", "crs": CRS84, "features": feats}


def write_geojson(fc: Dict, path: str) -> None:
    name = os.path.splitext(os.path.basename(path))[0]
    out = {"type": "FeatureCollection", "name": name, "crs": fc.get("crs", CRS84), "features": fc["features"]}
    tmp = path + ".tmp"
    with open(tmp, "w") as f:
        json.dump(out, f)
    os.replace(tmp, path)


def read_geojson(path: str) -> List[Dict]:
    """Point features back to ``{'lon','lat','confidence'}`` (``CheckpointManager.load_checkpoint``,
    ``_script/utils.py:108-117``)."""
    with open(path) as f:
        fc = json.load(f)
    out = []
    for ft in fc.get("features", []):
        x, y = ft["geometry"]["coordinates"][:2]
        out.append({'lon': x, 'lat': y, 'confidence': ft.get("properties", {}).get("confidence")})
    return out


# ---- _script/utils.py:68-146 -------------------------------------------------------------------------------
class CheckpointManager:
    def __init__(self, checkpoint_dir, prefix=''):
        self.checkpoint_dir = checkpoint_dir
        self.prefix = f"{prefix}_" if prefix else ""
        self.state_file = os.path.join(checkpoint_dir, f"{self.prefix}processing_state.json")
        self.data_file = os.path.join(checkpoint_dir, f"{self.prefix}latest_detections.geojson")
        self.temp_state_file = os.path.join(checkpoint_dir, f"{self.prefix}temp_state.json")
        self.temp_data_file = os.path.join(checkpoint_dir, f"{self.prefix}temp_detections.geojson")

    def save_checkpoint(self, processed_count, detections, total_tiles):
        """``processing_state.json`` (``processed_count``, ``total_tiles``, ``timestamp``, indent 2) and, when there
        are detections, ``latest_detections.geojson`` (``_script/utils.py:77-95``)."""
        state = {'processed_count': int(processed_count), 'total_tiles': int(total_tiles), 'timestamp': datetime.now().isoformat()}
        # detections first, then the state, each through a temporary file + os.replace: an interrupted save leaves the
        # previous (consistent) checkpoint in place instead of a truncated JSON that would silently restart from zero
        if detections:
            write_geojson(self._create_geodataframe(detections), self.data_file)
        with open(self.temp_state_file, 'w') as f:
            json.dump(state, f, indent=2)
        os.replace(self.temp_state_file, self.state_file)

    def load_checkpoint(self):
        """(processed_count, detections); a missing or unreadable checkpoint is ``(0, [])`` as in the reference
        (``_script/utils.py:97-124``)."""
        try:
            processed, dets = 0, []
            if os.path.exists(self.state_file):
                with open(self.state_file) as f:
                    processed = json.load(f)['processed_count']
            if os.path.exists(self.data_file):
                dets = read_geojson(self.data_file)
            return processed, dets
        except Exception as e:      # the reference prints and restarts from zero
            print(f"Error loading checkpoint: {e}")
            return 0, []

    def _create_geodataframe(self, detections):
        return create_geodataframe([d for d in detections if isinstance(d, dict) and 'lon' in d])


# ---- _script/utils.py:181-292 ------------------------------------------------------------------------------
class ResultsManager:
    def __init__(self, output_dir, prefix="detections", duplicate_distance=0, engine=None, shapefile=False):
        self.duplicate_distance = duplicate_distance      # metres
        self.shapefile = shapefile
        self.output_dir = output_dir
        self.output_file = os.path.join(output_dir, f"{prefix}_results.geojson")
        self.engine = engine                              # the B200 engine that runs the projection + dedup kernels
        os.makedirs(output_dir, exist_ok=True)

    def process_results(self, detections):
        """Final dedup, then ``<prefix>_results.geojson`` (``_script/utils.py:191-210``)."""
        if not detections:
            return []
        unique = self.remove_duplicates(detections)
        fc = create_geodataframe(unique)
        if fc["features"]:
            write_geojson(fc, self.output_file)
            if self.shapefile:           # opt-in second copy of the results as an ESRI shapefile (north_star: "shapefile output")
                write_shapefile(fc, os.path.splitext(self.output_file)[0] + ".shp")
        return unique

    def remove_duplicates(self, detections):
        """``ResultsManager.remove_duplicates`` (``_script/utils.py:212-274``): project to the UTM zone of the
        *mean* longitude (northern), walk the detections in descending confidence and drop every later one that
        lies strictly closer than ``duplicate_distance`` to a kept one (``<``, so the default distance 0 removes
        nothing), then project the survivors back -- their coordinates come out of a UTM round trip, as in the
        reference.  The projection and the greedy pass run on the device; survivors are returned in
        descending-confidence order (ties in input order; the reference's pandas sort leaves ties unspecified)."""
        if not detections:
            return []
        dets = [d for d in detections if isinstance(d, dict) and 'lon' in d and 'lat' in d]
        if not dets:
            return []
        lon = np.array([d['lon'] for d in dets], dtype=np.float64)
        lat = np.array([d['lat'] for d in dets], dtype=np.float64)
        conf = [d.get('confidence', 0.0) for d in dets]
        zone = geo.utm_zone_of(float(lon.mean()))
        order = sorted(range(len(dets)), key=lambda i: conf[i], reverse=True)
        if self.engine is None:
            raise RuntimeError("ResultsManager needs the B200 engine (engine=...) for its projection and dedup kernels; "
                               "there is no CPU fallback")
        import torch
        eng = self.engine
        rank = np.empty(len(dets), dtype=np.int64)
        rank[order] = np.arange(len(dets), dtype=np.int64)        # priority = position in the descending sort (int64 tie-break, constant conf)
        x, y = eng.utm_forward(torch.from_numpy(lon).to(eng.device), torch.from_numpy(lat).to(eng.device), zone, True)
        if self.duplicate_distance > 0:
            keep = eng.dedup(x, y, torch.zeros(len(dets), dtype=torch.float32, device=eng.device), float(self.duplicate_distance),
                             inclusive=False, tiebreak=torch.from_numpy(rank).to(eng.device)).cpu().numpy()
        else:
            keep = np.ones(len(dets), dtype=np.uint8)         # `distance < 0` is never true (:255): nothing is removed
        xs, ys = x.cpu().numpy(), y.cpu().numpy()
        idx = [i for i in order if keep[i]]
        blon, blat = geo.utm_inverse(xs[idx], ys[idx], zone, True)
        return [{'lon': float(a), 'lat': float(b), 'confidence': conf[i]} for a, b, i in zip(blon, blat, idx)]

    def save_intermediate_results(self, detections, processed_count, total_tiles):
        """``intermediate_results_<p>percent.geojson`` (``_script/utils.py:276-292``)."""
        if not detections:
            return
        pct = (processed_count / total_tiles) * 100
        fc = create_geodataframe(detections)
        if fc["features"]:
            write_geojson(fc, os.path.join(self.output_dir, f"intermediate_results_{pct:.1f}percent.geojson"))
