"""The reference's map-making loop (prcp_catalonia/5_prcp_prediction.R:228-313) on its REAL
inputs, GeoTIFF to GeoTIFF: the 87 Catalonian stations (temp_data/stfdf_temp.rda), the DEM grid
(dem_twi/dem_cat.tif, 50 046 valid cells), the daily tmin / tmax / IMERG rasters of
2016-01-01..04 -- packed into tests/golden/catalonia_maps.npz by tests/golden/make_catalonia.py.

The model (`../models/RFSI.rda`, 4_..._RK.R:1721-1729) is one of the reference's missing blobs, so it
is re-fitted here the way the script fits it (all days, locations == observations, n.obs = 7,
mtry 4, min.node.size 6, quantreg) with the host grower.  Everything after that is the hot
path: rasters decoded by librfsi_io.so, sampled / repaired / near.obs'd / predicted on the
device in ONE call, Int16 products written by librfsi_io.so -- and checked bit for bit against
the CPU oracle driven with host-sampled covariates."""
from pathlib import Path

import numpy as np
import pytest

import oracle
import rfsi_b200 as rb
from rfsi_b200 import geotiff, grow, mapping, proj
from tests.test_gpu_raster import np_extract

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
N_OBS = 7
COVS = ["imerg", "tmax", "tmin"]            # covariates <- c("prcp", "imerg","tmax","tmin") minus the response (4_..._RK.R:71)


def _table(g, name):
    v = g[name + "_x10"]
    return np.where(v == -32768, np.nan, v.astype(np.float64) / 10)


def _write_inputs(g, root):
    keys = geotiff.GeoKeys(g["dem_geokeys"], g["dem_geodoubles"], g["dem_geoascii"].tobytes())
    x, y, dx, dy = g["dem_geo"]
    geotiff.write_geotiff(root / "dem_cat.tif", g["dem"], x, y, dx, dy, nodata=-32767, geokeys=keys)
    days = [str(d) for d in g["days"]]
    x, y, dx, dy = g["t_geo"]
    for sub, key in (("min", "tmin"), ("max", "tmax")):
        (root / sub).mkdir()
        for k, d in enumerate(days):
            geotiff.write_geotiff(root / sub / f"{d}.tif", g[key][k], x, y, dx, dy, nodata=-32767, epsg=4326)
    (root / "imerg").mkdir()
    x, y, dx, dy = g["imerg_geo"]
    for k, d in enumerate(days):
        geotiff.write_geotiff(root / "imerg" / f"3B-HHR-L.MS.MRG.3IMERG.{d}-S233000-E235959.1410.V05B.1day.tif", g["imerg"][k],
                              x, y, dx, dy, epsg=4326, compression="none")
    return days


def test_catalonia_maps_from_geotiffs_to_geotiffs(tmp_path):
    g = np.load(GOLD / "catalonia_maps.npz")
    days = _write_inputs(g, tmp_path)
    prcp = _table(g, "prcp")
    tabs = {k: _table(g, k) for k in COVS}
    D, S = prcp.shape
    assert (D, S) == (1095, 87) and int(np.isnan(prcp).sum()) == 2945

    # gg / gg2 and the stations in UTM31 (5_prcp_prediction.R:38-46,236)
    grid = mapping.grid_from_dem(tmp_path / "dem_cat.tif", utm_zone=31)
    assert grid.npix == 50046 and (grid.info.ncol, grid.info.nrow) == (380, 280)
    st_xy = np.column_stack(proj.lonlat_to_utm(g["lonlat"][:, 0], g["lonlat"][:, 1], 31))

    # the model, fitted like 4_..._RK.R:1693-1727: every day, locations == observations
    dist, obs = rb.near_obs_days(st_xy, prcp, N_OBS)
    rows = ~np.isnan(prcp)
    cols = [tabs[k][rows] for k in COVS] + [dist[:, :, j][rows] for j in range(N_OBS)] + [obs[:, :, j][rows] for j in range(N_OBS)]
    X = np.stack(cols, axis=1)
    keep = ~np.isnan(X).any(axis=1)
    names = COVS + [f"dist{j + 1}" for j in range(N_OBS)] + [f"obs{j + 1}" for j in range(N_OBS)]
    assert keep.sum() == 92320                                # the reference's CV vectors have this length
    flat = grow.grow_forest(X[keep], prcp[rows][keep], names, num_trees=250, mtry=4, min_node_size=6, seed=42, quantreg=True)
    model = rb.RangerForest(flat)

    # the loop, all four days in one call
    cov = {"tmin": mapping.CovariateFiles([tmp_path / "min" / f"{d}.tif" for d in days], scale=10.0, divide=True),
           "tmax": mapping.CovariateFiles([tmp_path / "max" / f"{d}.tif" for d in days], scale=10.0, divide=True),
           "imerg": mapping.CovariateFiles([next((tmp_path / "imerg").glob(f"*{d}*.tif")) for d in days])}
    before = rb.launch_count()
    res = mapping.predict_maps(model, grid, st_xy, prcp[:4], days, N_OBS, cov, cov_fix=[("tmin", "tmax", 1.0)],
                               out_dir=tmp_path / "predictions" / "rfsi")
    assert rb.launch_count() > before
    assert res.days == days and res.pred_i16.shape == (4, 50046) and res.iqr_i16.shape == (4, 50046)
    assert (res.pred_i16 >= 0).all() and (res.iqr_i16 >= 0).all()            # precipitation, no NA inside the mask

    # oracle on a sample of pixels, covariates sampled on the host from the arrays in the fixture
    rng = np.random.default_rng(0)
    sel = np.sort(rng.choice(grid.npix, 2000, replace=False))
    ll = grid.xy_raster[sel]
    tg, ig = g["t_geo"], g["imerg_geo"]
    host = {"tmin": np.stack([np_extract(g["tmin"], tg[0], tg[1], tg[2], tg[3], -32767, 10.0, 0.0, ll, d, True) for d in range(4)]),
            "tmax": np.stack([np_extract(g["tmax"], tg[0], tg[1], tg[2], tg[3], -32767, 10.0, 0.0, ll, d, True) for d in range(4)]),
            "imerg": np.stack([np_extract(g["imerg"], ig[0], ig[1], ig[2], ig[3], np.nan, 1.0, 0.0, ll, d) for d in range(4)])}
    assert not any(np.isnan(v).any() for v in host.values())
    repaired = int((host["tmin"] > host["tmax"]).sum())
    host["tmin"] = np.where(host["tmin"] > host["tmax"], host["tmax"] - 1, host["tmin"])
    ref = oracle.predict_fused(flat.as_dict(), names, loc_xy=grid.xy[sel], obs_xy=st_xy, obs_z=prcp[:4], n_obs=N_OBS,
                               covariates=host, quantiles=[0.25, 0.5, 0.75])
    np.testing.assert_array_equal(res.pred_i16[:, sel].ravel(), oracle.centi_int16(ref["mean"].ravel()))
    q = ref["quantiles"]
    np.testing.assert_array_equal(res.iqr_i16[:, sel].ravel(), oracle.centi_int16(((q[2] - q[0]) / 2).ravel()))
    print(f"catalonia maps: 4 days x {grid.npix} pixels, {repaired} of {4 * len(sel)} sampled pixel-days had tmin > tmax, "
          f"mean prcp per day (mm): {np.round(res.pred_i16.mean(axis=1) / 100, 2).tolist()}")

    # the products: <day>.tif and <day>_iqr.tif, INT2S, NAflag -32767, LZW, on the DEM's grid and CRS
    assert [f.name for f in res.files] == [n for d in days for n in (f"{d}.tif", f"{d}_iqr.tif")]
    valid = (g["dem"] != -32767).ravel()
    for k, d in enumerate(days):
        for suffix, arr in (("", res.pred_i16[k]), ("_iqr", res.iqr_i16[k])):
            r, info = geotiff.read_geotiff(tmp_path / "predictions" / "rfsi" / f"{d}{suffix}.tif")
            assert (info.dtype, info.nodata, info.compression) == (np.int16, -32767.0, 5)
            assert (info.x_ul, info.y_ul, info.dx, info.dy) == tuple(g["dem_geo"])
            np.testing.assert_array_equal(info.geokeys.directory, g["dem_geokeys"])
            flat_r = r.data.ravel()
            np.testing.assert_array_equal(flat_r[valid], arr)
            assert (flat_r[~valid] == -32767).all()
    model.close()


def test_days_with_fewer_than_two_stations_are_skipped(tmp_path):
    """`if (length(obs_st) < 2) return(NULL)` (5_prcp_prediction.R:249-251)."""
    g = np.load(GOLD / "catalonia_maps.npz")
    days = _write_inputs(g, tmp_path)
    grid = mapping.grid_from_dem(tmp_path / "dem_cat.tif", utm_zone=31)
    st_xy = np.column_stack(proj.lonlat_to_utm(g["lonlat"][:, 0], g["lonlat"][:, 1], 31))
    prcp = _table(g, "prcp")[:4].copy()
    prcp[1, 1:] = np.nan                                      # one station left on day 2
    rng = np.random.default_rng(1)
    names = COVS + ["dist1", "dist2", "obs1", "obs2"]
    X = rng.uniform(0, 30, size=(500, len(names)))
    flat = grow.grow_forest(X, X[:, 5] + 0.1 * X[:, 1], names, num_trees=20, quantreg=True)
    model = rb.RangerForest(flat)
    cov = {"tmin": mapping.CovariateFiles([tmp_path / "min" / f"{d}.tif" for d in days], scale=10.0, divide=True),
           "tmax": mapping.CovariateFiles([tmp_path / "max" / f"{d}.tif" for d in days], scale=10.0, divide=True),
           "imerg": mapping.CovariateFiles([next((tmp_path / "imerg").glob(f"*{d}*.tif")) for d in days])}
    res = mapping.predict_maps(model, grid, st_xy, prcp, days, 2, cov, out_dir=tmp_path / "out")
    assert res.days == [days[0], days[2], days[3]] and res.pred_i16.shape == (3, grid.npix)
    assert sorted(f.name for f in res.files) == sorted(n for d in res.days for n in (f"{d}.tif", f"{d}_iqr.tif"))
    # day 3's map is what a single-day call gives (the right raster layer was paired with the right day)
    one = mapping.predict_maps(model, grid, st_xy, prcp[2:3], [days[2]], 2,
                               {k: mapping.CovariateFiles([c.paths[2]], c.scale, c.offset, c.divide) for k, c in cov.items()})
    np.testing.assert_array_equal(one.pred_i16[0], res.pred_i16[1])
    model.close()
