"""lon/lat -> UTM (gg2 <- spTransform(gg, utm31), prcp_catalonia/5_prcp_prediction.R:46).
PROJ cannot run here; anchors: the published CN Tower example of the UTM system, the meridian
arc, the symmetries of the projection, and the reference's own Croatian station table, which
carries UTM33 coordinates next to the Lat/Lon of the 1 km cell each station falls in."""
from pathlib import Path

import numpy as np

from rfsi_b200.proj import lonlat_to_utm, utm_zone

GOLD = Path(__file__).parent / "golden"


def test_known_points():
    # CN Tower, 43 38'33.24"N 79 23'13.7"W -> zone 17, 630 084 mE 4 833 438 mN (the textbook UTM example)
    e, n = lonlat_to_utm(-(79 + 23 / 60 + 13.7 / 3600), 43 + 38 / 60 + 33.24 / 3600, 17)
    assert abs(e - 630084) < 1.0 and abs(n - 4833438) < 1.0
    # on the central meridian: easting 500 km exactly, northing = k0 * meridian arc (45 deg: 4 984 944.378 m)
    e, n = lonlat_to_utm(3.0, 45.0, 31)
    assert e == 500000.0 and abs(n - 0.9996 * 4984944.378) < 1e-3
    assert lonlat_to_utm(3.0, 0.0, 31) == (500000.0, 0.0)
    # equator, 3 degrees off the central meridian (zone edge): 833 978.557 mE (the well-known zone half-width)
    e, n = lonlat_to_utm(6.0, 0.0, 31)
    assert abs(e - 833978.557) < 1e-2 and n == 0.0
    assert utm_zone(2.0) == 31 and utm_zone(16.4) == 33 and utm_zone(-79.4) == 17


def test_symmetries_and_south():
    lon = np.array([0.2, 1.7, 3.3]); lat = np.array([40.6, 41.9, 42.8])
    e, n = lonlat_to_utm(lon, lat, 31)
    e2, n2 = lonlat_to_utm(6.0 - lon, lat, 31)               # mirror about the central meridian
    np.testing.assert_allclose(e - 500000.0, 500000.0 - e2, rtol=0, atol=1e-8)
    np.testing.assert_allclose(n, n2, rtol=0, atol=1e-8)
    es, ns = lonlat_to_utm(lon, -lat, 31, south=True)
    np.testing.assert_allclose(es, e, atol=1e-8)
    np.testing.assert_allclose(ns, 10000000.0 - n, atol=1e-8)


def test_croatian_station_table():
    """temp_croatia/temp_data/stfdf.rda: UTM33 coordinates + the Lat/Lon covariates (values of the
    1 km cell around each station): the projected cell centre must lie within half a cell
    diagonal of the station, with no systematic shift."""
    g = np.load(GOLD / "croatia_llocv.npz")
    e, n = lonlat_to_utm(g["Lon"].astype(np.float64), g["Lat"].astype(np.float64), 33)
    de, dn = e - g["xy"][:, 0], n - g["xy"][:, 1]
    assert np.abs(de).max() < 520 and np.abs(dn).max() < 520
    assert abs(de.mean()) < 60 and abs(dn.mean()) < 60
