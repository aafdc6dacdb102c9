"""Fused near.obs -> forest over pixels x days on a real B200 vs the CPU oracle's
per-day filter + near.obs + cbind + predict loop (2_predictions.R:73-183)."""
import os

import numpy as np
import pytest

import oracle
import rfsi_b200 as rb
from rfsi_b200 import synth
from tests.util import same_mean

pytestmark = pytest.mark.gpu
PROBS = [0.25, 0.5, 0.75]


def _forest_for(case, T=60, quantreg=True, max_rows=4000):
    X, y = synth.training_table(case, lambda l, o, z, n: oracle.near_obs(l, o, z, n)[:2], max_rows=max_rows)
    return synth.grow_forest_sklearn(X, y, case.feature_names, num_trees=T, quantreg=quantreg, n_jobs=8)


def _check(got, ref, quant=True):
    same_mean(got["mean"], ref["mean"])
    ok = ~np.isnan(ref["mean"])
    if quant:
        np.testing.assert_array_equal(got["quantiles"][:, ok], ref["quantiles"][:, ok])
        assert np.isnan(got["quantiles"][:, ~ok]).all()


def test_c1_lattice_two_days_with_missing_stations():
    case = synth.make_case("c1", ndays=3)
    case.obs_z[1, ::5] = np.nan
    case.obs_z[2, 1::3] = np.nan
    flat = _forest_for(case)
    model = rb.RangerForest(flat)
    got = rb.predict_fused(model, obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, grid=case.grid,
                           npix=case.npix, quantiles=PROBS, centi_int16=True)
    ref = oracle.predict_fused(flat.as_dict(), case.feature_names, loc_xy=case.locations(), obs_xy=case.obs_xy,
                               obs_z=case.obs_z, n_obs=case.n_obs, quantiles=PROBS)
    _check(got, ref)
    # product epilogue: round-half-even(pred*100) Int16, iqr = round((q75-q25)/2*100)
    np.testing.assert_array_equal(got["mean_i16"].ravel(), oracle.centi_int16(ref["mean"].ravel()))
    iqr = (ref["quantiles"][2] - ref["quantiles"][0]) / 2
    np.testing.assert_array_equal(got["iqr_i16"].ravel(), oracle.centi_int16(iqr.ravel()))
    # explicit coordinates give the same answer as the raster form
    got2 = rb.predict_fused(model, obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs,
                            locations=case.locations(), quantiles=PROBS)
    np.testing.assert_array_equal(got2["mean"], got["mean"])
    np.testing.assert_array_equal(got2["quantiles"], got["quantiles"])


@pytest.mark.parametrize("name,rows", [("c3", 60), ("c4", 90)])
def test_real_shapes_with_covariates(name, rows):
    """temp_croatia (8 covariates, n=10, F=28) and prcp_catalonia (3 daily covariates, n=7,
    F=17) shapes on a band of raster rows starting mid-raster."""
    case = synth.make_case(name, ndays=4)
    flat = _forest_for(case, T=50)
    model = rb.RangerForest(flat)
    begin = 100 * case.ncol
    npix = rows * case.ncol
    loc = case.locations(begin, npix)
    cov = case.covariates(loc)
    mask = None
    if name == "c4":
        rng = np.random.default_rng(0)
        mask = (rng.uniform(size=npix) < 0.47).astype(np.uint8)     # dem_cat.tif: 47 % valid cells
    got = rb.predict_fused(model, obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, grid=case.grid,
                           pix_begin=begin, npix=npix, covariates=cov, quantiles=PROBS, pix_mask=mask)
    ref = oracle.predict_fused(flat.as_dict(), case.feature_names, loc_xy=loc, obs_xy=case.obs_xy,
                               obs_z=case.obs_z, n_obs=case.n_obs, covariates=cov, quantiles=PROBS, kdtree=True,
                               threads=8)
    if mask is not None:
        ref["mean"][:, mask == 0] = np.nan
        ref["quantiles"][:, :, mask == 0] = np.nan
    _check(got, ref)


def test_fallback_path_when_candidate_lists_run_dry(monkeypatch):
    """More stations missing on a day than the candidate lists have spares: those
    pixel-days are redone by the full-scan kernel and must still match."""
    case = synth.make_case("c1", ndays=2)
    rng = np.random.default_rng(3)
    case.obs_z[1, rng.choice(200, 120, replace=False)] = np.nan      # 60 % missing on day 1
    flat = _forest_for(case, T=30)
    model = rb.RangerForest(flat)
    got = rb.predict_fused(model, obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, grid=case.grid,
                           npix=case.npix, quantiles=PROBS)
    ref = oracle.predict_fused(flat.as_dict(), case.feature_names, loc_xy=case.locations(), obs_xy=case.obs_xy,
                               obs_z=case.obs_z, n_obs=case.n_obs, quantiles=PROBS)
    _check(got, ref)


def test_rank_caches_do_not_change_a_bit(monkeypatch):
    """The rank caches of the fused prologue (distance ranks of the candidates once per pixel, observation ranks
    once per launch; csrc/fused.cuh) against the search per pixel-day: days with 0, 1, 2 and many skipped
    candidates in front of a column (missing stations next to the pixel, the pixel's own station), a one-day
    call (no candidate cache by design), more stations than pixels (no observation table by design)."""
    case = synth.make_case("c1", ndays=4)
    loc = case.locations()
    d2 = ((loc[:, None, :] - case.obs_xy[None, :, :]) ** 2).sum(-1)
    order = np.argsort(d2[case.npix // 2 + 17])
    case.obs_z[1, order[0]] = np.nan                      # the nearest station of one pixel: shift 1 around it
    case.obs_z[2, order[:2]] = np.nan                     # shift 2
    case.obs_z[3, order[:5]] = np.nan                     # shift 5: beyond the cached shifts, searched
    case.obs_z[3, ::4] = np.nan
    flat = _forest_for(case)
    model = rb.RangerForest(flat)

    def run(**kw):
        return rb.predict_fused(model, obs_xy=case.obs_xy, n_obs=case.n_obs, grid=case.grid, quantiles=PROBS, **kw)

    ref = oracle.predict_fused(flat.as_dict(), case.feature_names, loc_xy=loc, obs_xy=case.obs_xy, obs_z=case.obs_z,
                               n_obs=case.n_obs, quantiles=PROBS)
    with_cache = run(obs_z=case.obs_z, npix=case.npix)
    _check(with_cache, ref)
    one_day = run(obs_z=case.obs_z[3:4], npix=case.npix)
    np.testing.assert_array_equal(one_day["mean"][0], with_cache["mean"][3])
    few = run(obs_z=case.obs_z, npix=40)                  # 200 stations, 40 pixels
    np.testing.assert_array_equal(few["mean"], with_cache["mean"][:, :40])
    monkeypatch.setenv("RFSI_B200_RANK_CACHE", "0")
    without = run(obs_z=case.obs_z, npix=case.npix)
    np.testing.assert_array_equal(without["mean"], with_cache["mean"])
    np.testing.assert_array_equal(without["quantiles"], with_cache["quantiles"])


def test_too_few_valid_stations_gives_na_and_warning():
    case = synth.make_case("c1", ndays=2)
    case.obs_z[1, 5:] = np.nan                                       # 5 valid stations, n.obs = 6
    flat = _forest_for(case, T=20, quantreg=False)
    model = rb.RangerForest(flat)
    with pytest.warns(UserWarning):
        got = rb.predict_fused(model, obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, grid=case.grid,
                               npix=case.npix)
    assert np.isnan(got["mean"][1]).all() and not np.isnan(got["mean"][0]).any()
    ref = oracle.predict_fused(flat.as_dict(), case.feature_names, loc_xy=case.locations(), obs_xy=case.obs_xy,
                               obs_z=case.obs_z, n_obs=case.n_obs)
    _check(got, ref, quant=False)


def test_na_covariate_is_reported_like_ranger():
    case = synth.make_case("c4", ndays=1)
    flat = _forest_for(case, T=10, quantreg=False)
    model = rb.RangerForest(flat)
    npix = 5 * case.ncol
    loc = case.locations(0, npix)
    cov = case.covariates(loc)
    cov["tmax"][0, 17] = np.nan
    with pytest.raises(rb.RfsiError) as e:
        rb.predict_fused(model, obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, locations=loc, covariates=cov)
    assert e.value.code == -4


def test_multi_chunk_multi_daybatch_host_pipeline():
    """> 2^20 pixels and > 16 days: exercises pixel chunking, day batches, both stream
    slots and candidate-list reuse; verified on a sample + invariants."""
    case = synth.make_case("c1s", ndays=18)
    rng = np.random.default_rng(1)
    for d in range(18):
        case.obs_z[d, rng.choice(1000, 30, replace=False)] = np.nan
    flat = _forest_for(case, T=16, quantreg=False, max_rows=3000)
    model = rb.RangerForest(flat)
    ncol, nrow = 1100, 1000                                          # 1.1 M cells (> one chunk)
    grid = (0.5, 0.5, 500.0 / ncol, 500.0 / nrow, ncol)
    got = rb.predict_fused(model, obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, grid=grid,
                           npix=ncol * nrow)
    assert got["mean"].shape == (18, ncol * nrow) and not np.isnan(got["mean"]).any()
    sel = np.sort(rng.choice(ncol * nrow, 3000, replace=False))
    loc = np.stack([grid[0] + (sel % ncol) * grid[2], grid[1] + (sel // ncol) * grid[3]], 1)
    ref = oracle.predict_fused(flat.as_dict(), case.feature_names, loc_xy=loc, obs_xy=case.obs_xy,
                               obs_z=case.obs_z, n_obs=case.n_obs, kdtree=True, threads=8)
    same_mean(got["mean"][:, sel], ref["mean"])


@pytest.mark.parametrize("T", [129, 200, 256, 257, 300, 512])
def test_quantile_products_for_129_to_512_trees(T):
    """129..512 trees (8 or 16 keys per lane in the quantile kernel, both sides of the 256 boundary) through the fused call:
    rows of the fallback list (60 % of the stations missing on day 1), masked pixels and a day with too few stations.
    Seven quantile levels = three passes of three in the one-pass selection; the Int16 products come from the same selection."""
    case = synth.make_case("c4", ndays=4)
    rng = np.random.default_rng(T)
    case.obs_z[1, rng.choice(case.obs_z.shape[1], int(0.6 * case.obs_z.shape[1]), replace=False)] = np.nan
    case.obs_z[3, case.n_obs - 1:] = np.nan                          # fewer valid stations than n.obs: the whole day is NA
    flat = _forest_for(case, T=T, max_rows=1500)
    if T == 300:                                                     # heavy ties: a handful of distinct leaf samples
        rnv = flat.random_node_values.copy(); ok = ~np.isnan(rnv); rnv[ok] = np.round(rnv[ok], 0); flat.random_node_values = rnv
    model = rb.RangerForest(flat)
    begin, npix = 120 * case.ncol, 40 * case.ncol
    loc = case.locations(begin, npix)
    cov = {k: np.array(v, copy=True) for k, v in case.covariates(loc).items()}
    mask = (rng.uniform(size=npix) < 0.7).astype(np.uint8)
    probs = [0.0, 0.1, 0.25, 0.5, 0.75, 0.9, 1.0]
    with pytest.warns(UserWarning):
        got = rb.predict_fused(model, obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, grid=case.grid, pix_begin=begin,
                               npix=npix, covariates=cov, quantiles=probs, pix_mask=mask, centi_int16=True)
    ref = oracle.predict_fused(flat.as_dict(), case.feature_names, loc_xy=loc, obs_xy=case.obs_xy, obs_z=case.obs_z,
                               n_obs=case.n_obs, covariates=cov, quantiles=probs, kdtree=True, threads=8)
    ref["mean"][:, mask == 0] = np.nan
    ref["quantiles"][:, :, mask == 0] = np.nan
    assert np.isnan(ref["mean"][3]).all() and not np.isnan(ref["mean"][1][mask == 1]).any()
    same_mean(got["mean"], ref["mean"])
    np.testing.assert_array_equal(got["quantiles"], ref["quantiles"])            # NaN == NaN here
    iqr = (ref["quantiles"][-1] - ref["quantiles"][0]) / 2
    np.testing.assert_array_equal(got["iqr_i16"].ravel(), oracle.centi_int16(iqr.ravel()))
