"""Covariate sampling on the device (raster::extract, method 'simple') and its use inside the
fused predictor, against a NumPy nearest-cell restatement
(prcp_catalonia/5_prcp_prediction.R:253-276)."""
import numpy as np
import pytest

import oracle
import rfsi_b200 as rb
from oracle import rfsi_oracle_np as onp
from rfsi_b200 import grow, synth
from tests.util import same_mean

pytestmark = pytest.mark.gpu


def np_extract(data, x_ul, y_ul, dx, dy, nodata, scale, offset, xy, layer=0, divide=False):
    return onp.raster_extract(data, x_ul, y_ul, dx, dy, nodata, xy, layer, scale, offset, divide)


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


def test_divide_is_a_division_not_a_multiplication_by_the_reciprocal():
    """extract(tmin, gg) / 10 (5_prcp_prediction.R:254): x/10 and x*0.1 differ in the last bit for
    about a third of the integers; the engine must give R's value."""
    data = np.arange(-400, 400, dtype=np.int16).reshape(20, 40)
    xy = np.column_stack([np.tile(np.arange(40) + 0.5, 20), np.repeat(20 - (np.arange(20) + 0.5), 40)])
    r = rb.Raster(data, 0.0, 20.0, 1.0, 1.0, nodata=-32767, scale=10.0, divide=True)
    got = rb.raster_extract(r, locations=xy)
    np.testing.assert_array_equal(got, data.ravel().astype(np.float64) / 10)
    mul = rb.raster_extract(rb.Raster(data, 0.0, 20.0, 1.0, 1.0, nodata=-32767, scale=0.1), locations=xy)
    assert (mul != got).sum() > 100                       # the two are not interchangeable
    with pytest.raises(rb.RfsiError):
        rb.raster_extract(rb.Raster(data, 0.0, 20.0, 1.0, 1.0, scale=0.0, divide=True), locations=xy)


def test_fused_covariate_repair_and_separate_sampling_coordinates():
    """5_prcp_prediction.R:253-283: rasters sampled at `gg` (their own CRS), near.obs at `gg2`
    (projected), tmin repaired where it exceeds tmax -- all inside the one fused call."""
    case = synth.make_case("c4", ndays=3)
    ncol, nrow = case.ncol, case.nrow
    rng = np.random.default_rng(5)
    loc = case.locations()
    # the rasters live in another coordinate system: an affine image of the prediction coordinates
    cov_loc = np.column_stack([0.16 + loc[:, 0] * 0.008, 40.5 + loc[:, 1] * 0.008])
    geo = dict(x_ul=0.16, y_ul=40.5 + nrow * 0.008, dx=0.008, dy=0.008)
    tmax = rng.integers(-50, 350, size=(3, nrow, ncol)).astype(np.int16)
    tmin = (tmax - rng.integers(-30, 120, size=tmax.shape)).astype(np.int16)     # above tmax in ~20 % of the cells
    imerg = rng.integers(0, 500, size=(3, 30, 40)).astype(np.uint16)
    R = {"tmax": rb.Raster(tmax, nodata=-32767, scale=10.0, divide=True, **geo),
         "tmin": rb.Raster(tmin, nodata=-32767, scale=10.0, divide=True, **geo),
         "imerg": rb.Raster(imerg, 0.0, 40.5 + nrow * 0.008 + 0.1, 0.1, 0.1)}
    host = {k: np.stack([np_extract(r.data, r.x_ul, r.y_ul, r.dx, r.dy, r.nodata, r.scale, r.offset, cov_loc, d, r.divide)
                         for d in range(3)]) for k, r in R.items()}
    assert not any(np.isnan(v).any() for v in host.values())
    bad = host["tmin"] > host["tmax"]
    assert 0.05 < bad.mean() < 0.5
    fixed = dict(host)
    fixed["tmin"] = onp.cov_fix(host["tmin"], host["tmax"], 1.0)            # ifelse(tmin > tmax, tmax-1, tmin)

    def station_cov(k):                             # the training table samples the same rasters at the stations
        def f(xy, d):
            at = np.column_stack([0.16 + xy[:, 0] * 0.008, 40.5 + xy[:, 1] * 0.008])
            v = {n: np_extract(R[n].data, R[n].x_ul, R[n].y_ul, R[n].dx, R[n].dy, R[n].nodata, R[n].scale, R[n].offset,
                               at, d, R[n].divide) for n in R}
            v["tmin"] = np.where(v["tmin"] > v["tmax"], v["tmax"] - 1, v["tmin"])
            return v[k]
        return f
    case.cov_daily = {k: station_cov(k) for k in R}
    X, y = synth.training_table(case, lambda l, o, z, n: oracle.near_obs(l, o, z, n)[:2])
    flat = grow.grow_forest(X, y, case.feature_names, num_trees=40, quantreg=True)
    model = rb.RangerForest(flat)
    kw = dict(obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, quantiles=[0.25, 0.5, 0.75])
    a = rb.predict_fused(model, locations=loc, cov_locations=cov_loc, covariates=R, cov_fix=[("tmin", "tmax", 1.0)], **kw)
    b = rb.predict_fused(model, locations=loc, covariates=fixed, **kw)
    np.testing.assert_array_equal(a["mean"], b["mean"])
    np.testing.assert_array_equal(a["quantiles"], b["quantiles"])
    unfixed = rb.predict_fused(model, locations=loc, cov_locations=cov_loc, covariates=R, **kw)
    assert (unfixed["mean"] != a["mean"]).any()            # the repair matters for this forest
    # repairs also apply to host-array covariates, static pairs included; mixed dynamics are refused
    c = rb.predict_fused(model, locations=loc, covariates=host, cov_fix=[("tmin", "tmax", 1.0)], **kw)
    np.testing.assert_array_equal(c["mean"], a["mean"])
    with pytest.raises(rb.RfsiError):
        rb.predict_fused(model, locations=loc, covariates={**host, "tmax": host["tmax"][0]}, cov_fix=[("tmin", "tmax", 1.0)], **kw)
    sel = np.sort(rng.choice(case.npix, 1200, replace=False))
    ref = oracle.predict_fused(flat.as_dict(), case.feature_names, loc_xy=loc[sel], obs_xy=case.obs_xy, obs_z=case.obs_z,
                               n_obs=case.n_obs, covariates={k: v[:, sel] for k, v in fixed.items()})
    same_mean(a["mean"][:, sel], ref["mean"])
