"""Longitude/latitude (WGS84) -> UTM, the one spTransform the mapping scripts need.

prcp_catalonia/5_prcp_prediction.R:44-46,236 keeps the prediction grid twice: `gg` in
lon/lat (the covariate rasters' CRS, used by raster::extract) and
`gg2 <- spTransform(gg, "+proj=utm +zone=31 +ellps=WGS84")` in metres (used by near.obs); the
stations go through the same transform.  PROJ is not part of this build; this is the
Krueger series for the transverse Mercator to n^6 (the expansion PROJ's `tmerc`/`utm`
implements: Karney 2011, "Transverse Mercator with an accuracy of a few nanometers"), host-side NumPy.
UNPINNED against PROJ (R/PROJ cannot run here); checked to the rounding of the stored
coordinates against the reference's own Croatian station table, whose UTM33 coordinates and
Lat/Lon columns are both in temp_croatia/temp_data/stfdf.rda (tests/test_proj.py).
"""
from __future__ import annotations

import numpy as np

WGS84_A = 6378137.0
WGS84_F = 1.0 / 298.257223563
K0 = 0.9996


def _alpha(n: float):
    n2, n3, n4, n5, n6 = n * n, n ** 3, n ** 4, n ** 5, n ** 6
    return (
        n / 2 - 2 * n2 / 3 + 5 * n3 / 16 + 41 * n4 / 180 - 127 * n5 / 288 + 7891 * n6 / 37800,
        13 * n2 / 48 - 3 * n3 / 5 + 557 * n4 / 1440 + 281 * n5 / 630 - 1983433 * n6 / 1935360,
        61 * n3 / 240 - 103 * n4 / 140 + 15061 * n5 / 26880 + 167603 * n6 / 181440,
        49561 * n4 / 161280 - 179 * n5 / 168 + 6601661 * n6 / 7257600,
        34729 * n5 / 80640 - 3418889 * n6 / 1995840,
        212378941 * n6 / 319334400,
    )


def utm_zone(lon) -> int:
    """The UTM zone a longitude falls in (no Norway/Svalbard exceptions)."""
    return int(np.floor((float(np.mean(lon)) + 180.0) / 6.0)) % 60 + 1


def lonlat_to_utm(lon, lat, zone: int, south: bool = False):
    """(lon, lat) in degrees -> (easting, northing) in metres, "+proj=utm +zone=<zone> +ellps=WGS84"."""
    lon = np.asarray(lon, dtype=np.float64)
    lat = np.asarray(lat, dtype=np.float64)
    f = WGS84_F
    n = f / (2.0 - f)
    e = np.sqrt(f * (2.0 - f))
    A = WGS84_A / (1.0 + n) * (1.0 + n ** 2 / 4 + n ** 4 / 64 + n ** 6 / 256)
    lam = np.deg2rad(lon - (6.0 * zone - 183.0))
    phi = np.deg2rad(lat)
    s = np.sin(phi)
    t = np.sinh(np.arctanh(s) - e * np.arctanh(e * s))      # tan of the conformal latitude
    xi = np.arctan2(t, np.cos(lam))
    eta = np.arcsinh(np.sin(lam) / np.hypot(t, np.cos(lam)))
    x, y = eta.copy(), xi.copy()
    for j, a in enumerate(_alpha(n), start=1):
        x += a * np.cos(2 * j * xi) * np.sinh(2 * j * eta)
        y += a * np.sin(2 * j * xi) * np.cosh(2 * j * eta)
    return 500000.0 + K0 * A * x, K0 * A * y + (10000000.0 if south else 0.0)
