"""Host side of the map-making loop (rfsi_b200/mapping.py): the prediction grid from a DEM
(`gg` / `gg2`, prcp_catalonia/5_prcp_prediction.R:38-46), stacking the daily covariate files,
putting products back on the grid.  No compute: the GPU part is tests/test_gpu_maps.py."""
from pathlib import Path

import numpy as np
import pytest

from rfsi_b200 import geotiff, mapping, proj

GOLD = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def cat(tmp_path_factory):
    g = np.load(GOLD / "catalonia_maps.npz")
    root = tmp_path_factory.mktemp("cat")
    keys = geotiff.GeoKeys(g["dem_geokeys"], g["dem_geodoubles"], g["dem_geoascii"].tobytes())
    x, y, dx, dy = g["dem_geo"]
    geotiff.write_geotiff(root / "dem_cat.tif", g["dem"], x, y, dx, dy, nodata=-32767, geokeys=keys)
    x, y, dx, dy = g["t_geo"]
    for k, d in enumerate(g["days"]):
        geotiff.write_geotiff(root / f"min_{d}.tif", g["tmin"][k], x, y, dx, dy, nodata=-32767, epsg=4326)
    return g, root


def test_grid_from_dem_is_the_valid_cells_in_both_crs(cat):
    g, root = cat
    grid = mapping.grid_from_dem(root / "dem_cat.tif", utm_zone=31)
    valid = g["dem"] != -32767
    assert grid.npix == 50046 == int(valid.sum())
    np.testing.assert_array_equal(grid.cell, np.flatnonzero(valid.ravel()))
    x0, y0, dx, dy = g["dem_geo"]
    row, col = np.divmod(grid.cell, 380)
    np.testing.assert_array_equal(grid.xy_raster[:, 0], x0 + (col + 0.5) * dx)       # cell centres, lon/lat
    np.testing.assert_array_equal(grid.xy_raster[:, 1], y0 - (row + 0.5) * dy)
    e, n = proj.lonlat_to_utm(grid.xy_raster[:, 0], grid.xy_raster[:, 1], 31)
    np.testing.assert_array_equal(grid.xy, np.column_stack([e, n]))
    # Catalonia in UTM31: eastings 260-530 km, northings 4.49-4.75 Mm; ~0.7 x 0.93 km cells
    assert 2.5e5 < grid.xy[:, 0].min() and grid.xy[:, 0].max() < 5.4e5
    assert 4.48e6 < grid.xy[:, 1].min() and grid.xy[:, 1].max() < 4.76e6
    same_row = np.flatnonzero(np.diff(grid.cell) == 1)[:1000]
    step = np.hypot(*(grid.xy[same_row + 1] - grid.xy[same_row]).T)
    assert 650 < step.min() and step.max() < 720
    # without a projection both coordinate sets are the same object (no second upload)
    plain = mapping.grid_from_dem(root / "dem_cat.tif")
    assert plain.xy is plain.xy_raster
    # products go back on the full grid with the NAflag elsewhere
    vals = np.arange(grid.npix, dtype=np.int64).astype(np.int16)
    full = grid.to_raster_i16(vals)
    assert full.shape == (280, 380) and full.dtype == np.int16
    np.testing.assert_array_equal(full[valid], vals)
    assert (full[~valid] == -32767).all()


def test_covariate_files_stack_daily_rasters_and_refuse_mismatches(cat, tmp_path):
    g, root = cat
    days = [str(d) for d in g["days"]]
    c = mapping.CovariateFiles([root / f"min_{d}.tif" for d in days], scale=10.0, divide=True)
    r = c.load(4)
    assert r.data.shape == (4, 280, 380) and r.data.dtype == np.int16 and r.nodata == -32767.0
    assert (r.scale, r.divide, r.dx, r.dy) == (10.0, True, g["t_geo"][2], g["t_geo"][3])
    np.testing.assert_array_equal(r.data, g["tmin"])
    one = mapping.CovariateFiles(root / "dem_cat.tif").load(4)         # a static covariate: one file
    assert one.data.shape == (280, 380) and one.data.dtype == np.float32
    with pytest.raises(ValueError):
        c.load(3)                                                      # four files for three days
    geotiff.write_geotiff(tmp_path / "other.tif", g["tmin"][0][:, :379], *g["t_geo"], nodata=-32767)
    with pytest.raises(ValueError):
        mapping.CovariateFiles([root / f"min_{days[0]}.tif", tmp_path / "other.tif"]).load(2)
    geotiff.write_geotiff(tmp_path / "nod.tif", g["tmin"][0], *g["t_geo"], nodata=-9999)
    with pytest.raises(ValueError):
        mapping.CovariateFiles([root / f"min_{days[0]}.tif", tmp_path / "nod.tif"]).load(2)
