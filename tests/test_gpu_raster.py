"""Covariate sampling on the device (raster::extract, method 'simple') and its use inside the
fused predictor, against a NumPy nearest-cell restatement
(prcp_catalonia/5_prcp_prediction.R:253-276)."""
import numpy as np
import pytest

import oracle
import rfsi_b200 as rb
from rfsi_b200 import grow, synth
from tests.util import same_mean

pytestmark = pytest.mark.gpu


def np_extract(data, x_ul, y_ul, dx, dy, nodata, scale, offset, xy, layer=0):
    d = data if data.ndim == 3 else data[None]
    nrow, ncol = d.shape[1:]
    x, y = xy[:, 0], xy[:, 1]
    inside = (x >= x_ul) & (x <= x_ul + ncol * dx) & (y <= y_ul) & (y >= y_ul - nrow * dy)
    c = np.minimum(np.floor((x - x_ul) / dx), ncol - 1).astype(np.int64)
    r = np.minimum(np.floor((y_ul - y) / dy), nrow - 1).astype(np.int64)
    out = np.full(len(x), np.nan)
    raw = d[layer][np.clip(r, 0, nrow - 1), np.clip(c, 0, ncol - 1)].astype(np.float64)
    ok = inside & ~(raw == nodata) & ~np.isnan(raw)
    out[ok] = raw[ok] * scale + offset
    return out


@pytest.mark.parametrize("dtype", ["float64", "float32", "int32", "int16", "uint16", "uint8"])
def test_extract_matches_nearest_cell_lookup(dtype):
    rng = np.random.default_rng(3)
    nrow, ncol = 37, 53
    data = rng.integers(0, 200, size=(3, nrow, ncol)).astype(dtype)
    data[:, 5, 7] = 99                                       # nodata cells
    r = rb.Raster(data, x_ul=100.0, y_ul=900.0, dx=2.5, dy=4.0, nodata=99, scale=0.1, offset=-3.0)
    xy = np.column_stack([rng.uniform(90, 100 + ncol * 2.5 + 10, 5000), rng.uniform(900 - nrow * 4 - 10, 910, 5000)])
    xy[:4] = [[100.0, 900.0], [100 + ncol * 2.5, 900 - nrow * 4.0], [100.0, 900 - nrow * 4.0], [102.5, 896.0]]  # edges
    for layer in (0, 2):
        got = rb.raster_extract(r, locations=xy, layer=layer)
        exp = np_extract(data, 100.0, 900.0, 2.5, 4.0, 99, 0.1, -3.0, xy, layer)
        np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
        np.testing.assert_array_equal(got[~np.isnan(exp)], exp[~np.isnan(exp)])
    # raster form of the locations: the raster's own cell centres return the raster itself
    got = rb.raster_extract(rb.Raster(data[1].astype(dtype), 100.0, 900.0, 2.5, 4.0),
                            grid=(100 + 1.25, 900 - 2.0, 2.5, -4.0, ncol), npix=nrow * ncol)
    np.testing.assert_array_equal(got, data[1].astype(np.float64).ravel())


def test_fused_with_raster_covariates_matches_host_sampled_covariates():
    """prcp_catalonia shape: tmax/tmin in tenths of a degree on the prediction grid (Int32, one
    layer per day), imerg on a coarser grid (UInt16), sampled on the device inside the fused call."""
    case = synth.make_case("c4", ndays=3)
    ncol, nrow = case.ncol, case.nrow
    rng = np.random.default_rng(0)
    loc = case.locations()
    tmax = np.stack([np.round(100 * (2 + case.cov_daily["tmax"](loc, d))).reshape(nrow, ncol)[::-1] for d in range(3)]).astype(np.int32)
    tmin = (tmax - rng.integers(10, 80, size=tmax.shape)).astype(np.int32)
    imerg = rng.integers(0, 500, size=(3, 28, 38)).astype(np.uint16)
    # grids: cell centres of `case` are (0.5 + c, 0.5 + r); north-up rasters have row 0 at the top
    R = {"tmax": rb.Raster(tmax, 0.0, float(nrow), 1.0, 1.0, nodata=-32767, scale=0.1),
         "tmin": rb.Raster(tmin, 0.0, float(nrow), 1.0, 1.0, nodata=-32767, scale=0.1),
         "imerg": rb.Raster(imerg, 0.0, float(nrow), 10.0, 10.0, scale=0.1)}
    host_cov = {k: np.stack([np_extract(r.data, r.x_ul, r.y_ul, r.dx, r.dy, r.nodata, r.scale, r.offset, loc, d)
                             for d in range(3)]) for k, r in R.items()}
    assert not np.isnan(host_cov["imerg"]).any()

    def train_no(l, o, z, n):
        return oracle.near_obs(l, o, z, n)[:2]
    case.cov_daily = {k: (lambda xy, d, k=k: np_extract(R[k].data, R[k].x_ul, R[k].y_ul, R[k].dx, R[k].dy, R[k].nodata,
                                                          R[k].scale, R[k].offset, xy, d)) for k in R}
    X, y = synth.training_table(case, train_no)
    flat = grow.grow_forest(X, y, case.feature_names, num_trees=40, quantreg=True)
    model = rb.RangerForest(flat)
    kw = dict(obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, grid=case.grid, npix=case.npix,
              quantiles=[0.25, 0.5, 0.75])
    a = rb.predict_fused(model, covariates=R, **kw)
    b = rb.predict_fused(model, covariates=host_cov, **kw)
    np.testing.assert_array_equal(a["mean"], b["mean"])
    np.testing.assert_array_equal(a["quantiles"], b["quantiles"])
    sel = np.sort(rng.choice(case.npix, 1500, replace=False))
    ref = oracle.predict_fused(flat.as_dict(), case.feature_names, loc_xy=loc[sel], obs_xy=case.obs_xy, obs_z=case.obs_z,
                               n_obs=case.n_obs, covariates={k: v[:, sel] for k, v in host_cov.items()})
    same_mean(a["mean"][:, sel], ref["mean"])
    # explicit locations + a mixed dict (one raster, two arrays)
    c = rb.predict_fused(model, obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, locations=loc[sel],
                         covariates={"imerg": R["imerg"], "tmax": host_cov["tmax"][:, sel], "tmin": host_cov["tmin"][:, sel]})
    np.testing.assert_array_equal(c["mean"], a["mean"][:, sel])
