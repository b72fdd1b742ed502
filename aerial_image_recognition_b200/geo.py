"""WGS84 <-> UTM in float64 without pyproj.

The reference projects through ``pyproj.Transformer`` (``_script/utils.py:36-41``,
``simple_detector.py:551-556``); pyproj is not installable here, so the transverse Mercator is
restated from its published form: the Krueger series in the third flattening ``n`` (Karney 2011,
"Transverse Mercator with an accuracy of a few nanometers", eqs. 35-36) to order n^4 -- the
same series PROJ's ``etmerc`` evaluates (to n^6) -- on the WGS84 ellipsoid, k0 = 0.9996, false
easting 500 km, false northing 10 000 km in the south.  Truncation error is below 1e-6 m inside a
UTM zone; results cannot be bit-compared with PROJ in this environment.

Host-side NumPy: tile lists and result files are host work in the reference too; the bulk
projection of detections for the dedup runs on the device (``b2d_utm_forward``).
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

A_WGS84 = 6378137.0
F_WGS84 = 1.0 / 298.257223563
K0 = 0.9996
_N = F_WGS84 / (2.0 - F_WGS84)
_AA = A_WGS84 / (1.0 + _N) * (1.0 + _N ** 2 / 4.0 + _N ** 4 / 64.0)
_ALPHA = (
    _N / 2 - 2 * _N ** 2 / 3 + 5 * _N ** 3 / 16 + 41 * _N ** 4 / 180,
    13 * _N ** 2 / 48 - 3 * _N ** 3 / 5 + 557 * _N ** 4 / 1440,
    61 * _N ** 3 / 240 - 103 * _N ** 4 / 140,
    49561 * _N ** 4 / 161280,
)
_BETA = (
    _N / 2 - 2 * _N ** 2 / 3 + 37 * _N ** 3 / 96 - _N ** 4 / 360,
    _N ** 2 / 48 + _N ** 3 / 15 - 437 * _N ** 4 / 1440,
    17 * _N ** 3 / 480 - 37 * _N ** 4 / 840,
    4397 * _N ** 4 / 161280,
)
_E = math.sqrt(F_WGS84 * (2.0 - F_WGS84))


def utm_zone_of(lon: float) -> int:
    """``int((lon + 180) / 6) + 1`` -- the reference's zone rule (``_script/utils.py:20``, ``simple_detector.py:546``)."""
    return int((lon + 180) / 6) + 1


def utm_epsg(lon: float, lat: float) -> str:
    """``TileGenerator.get_utm_epsg`` (``_script/utils.py:17-24``)."""
    epsg = 32600 + utm_zone_of(lon)
    if lat < 0:
        epsg += 100
    return f"EPSG:{epsg}"


def _lon0(zone: int) -> float:
    return math.radians((zone - 1) * 6 - 180 + 3)


def utm_forward(lon, lat, zone: int, north: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    lon = np.asarray(lon, np.float64); lat = np.asarray(lat, np.float64)
    phi = np.radians(lat); lam = np.radians(lon) - _lon0(zone)
    s = np.sin(phi)
    t = np.sinh(np.arctanh(s) - _E * np.arctanh(_E * s))
    xi = np.arctan2(t, np.cos(lam))
    eta = np.arctanh(np.sin(lam) / np.sqrt(1.0 + t * t))
    x = eta.copy(); y = xi.copy()
    for j, a in enumerate(_ALPHA, start=1):
        x = x + a * np.cos(2 * j * xi) * np.sinh(2 * j * eta)
        y = y + a * np.sin(2 * j * xi) * np.cosh(2 * j * eta)
    return 500000.0 + K0 * _AA * x, K0 * _AA * y + (0.0 if north else 10000000.0)


def utm_inverse(easting, northing, zone: int, north: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    e = np.asarray(easting, np.float64); nn = np.asarray(northing, np.float64)
    xi = (nn - (0.0 if north else 10000000.0)) / (K0 * _AA)
    eta = (e - 500000.0) / (K0 * _AA)
    xi_p = xi.copy(); eta_p = eta.copy()
    for j, b in enumerate(_BETA, start=1):
        xi_p = xi_p - b * np.sin(2 * j * xi) * np.cosh(2 * j * eta)
        eta_p = eta_p - b * np.cos(2 * j * xi) * np.sinh(2 * j * eta)
    tau_p = np.sin(xi_p) / np.sqrt(np.sinh(eta_p) ** 2 + np.cos(xi_p) ** 2)      # tan of the conformal latitude
    lam = np.arctan2(np.sinh(eta_p), np.cos(xi_p))
    tau = tau_p.copy()                                                            # Newton on tau = tan(phi), Karney eqs. 19-21
    for _ in range(5):
        sig = np.sinh(_E * np.arctanh(_E * tau / np.sqrt(1.0 + tau * tau)))
        f = tau * np.sqrt(1.0 + sig * sig) - sig * np.sqrt(1.0 + tau * tau) - tau_p
        df = (np.sqrt(1.0 + sig * sig) * np.sqrt(1.0 + tau * tau) - sig * tau) * (1.0 - _E * _E) * np.sqrt(1.0 + tau * tau) \
            / (1.0 + (1.0 - _E * _E) * tau * tau)
        tau = tau - f / df
    return np.degrees(lam + _lon0(zone)), np.degrees(np.arctan(tau))
