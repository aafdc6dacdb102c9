"""BASELINE.json configs at their full sizes on a real B200, checked through
size-independent properties plus oracle samples (the oracle cannot finish the full sizes in
seconds).  Tolerances: indices/distances bit-exact; means 1e-12 relative (bit-equal in the
default summation mode)."""
import numpy as np
import pytest

import oracle
import rfsi_b200 as rb
from rfsi_b200 import grow, synth
from tests.util import same_mean

pytestmark = pytest.mark.gpu


def _knn_arrays(df, n):
    a = df.to_numpy()
    return a[:, :n], a[:, n:]


def test_c2_sim500_full_grid_n40():
    """sim_500.R: 500x500 grid, 500 stations on cell centres, near.obs once with n.obs = 40,
    then forests on column subsets c(1:n, 41:(40+n)) (sim_500.R:131-152)."""
    case = synth.make_case("c2")
    loc = case.locations()
    df, idx = rb.near_obs(loc, np.column_stack([case.obs_xy, case.obs_z[0]]), zcol=2, n_obs=40, return_idx=True)
    d, o = _knn_arrays(df, 40)
    assert d.shape == (250_000, 40) and np.all(np.diff(d, axis=1) >= 0)
    assert (d[:, 0] > 0).all()                                          # self matches dropped
    rec = np.sqrt(((loc[:, None, :] - case.obs_xy[idx[:, [0, 17, 39]] - 1]) ** 2).sum(-1))
    np.testing.assert_array_equal(rec, d[:, [0, 17, 39]])               # dist_k == ||loc - obs[idx_k]||
    np.testing.assert_array_equal(case.obs_z[0][idx - 1], o)
    assert (np.diff(d, axis=1) == 0).mean() > 0.01                      # the lattice really has ties
    sel = np.random.default_rng(0).choice(250_000, 4000, replace=False)
    ed, eo, ei, _ = oracle.near_obs(loc[sel], case.obs_xy, case.obs_z[0], 40, kdtree=True, nthreads=8)
    np.testing.assert_array_equal(d[sel], ed); np.testing.assert_array_equal(idx[sel], ei)
    # a forest on the n = 5 column subset, predicted for the whole grid by column NAME
    n = 5
    cols = [f"dist{j}" for j in range(1, n + 1)] + [f"obs{j}" for j in range(1, n + 1)]
    tr = rb.near_obs(case.obs_xy, np.column_stack([case.obs_xy, case.obs_z[0]]), zcol=2, n_obs=40)
    flat = grow.grow_forest(tr[cols].to_numpy(), case.obs_z[0], cols, num_trees=250)
    pred = rb.predict(rb.RangerForest(flat), df).predictions            # df carries all 80 columns
    same_mean(pred[sel], oracle.forest_predict(flat.as_dict(), df[cols].to_numpy()[sel], threads=8))


@pytest.mark.parametrize("name", ["c3", "c4"])
def test_real_shapes_full_raster_all_mapped_days(name):
    """temp_croatia (490x487, 157 stations, n=10, F=28) and prcp_catalonia (380x280, 87
    stations, n=7, F=17) full rasters x the 4 mapped days (2_predictions.R:65;
    5_prcp_prediction.R:32): mean, quantiles (.25,.5,.75) and the Int16 IQR product."""
    case = synth.make_case(name, ndays=4)
    X, y = synth.training_table(case, lambda l, o, z, n: oracle.near_obs(l, o, z, n)[:2])
    flat = grow.grow_forest(X, y, case.feature_names, num_trees=250, min_node_size=6, quantreg=True)
    model = rb.RangerForest(flat)
    loc = case.locations()
    cov = case.covariates(loc)
    got = rb.predict_fused(model, obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, grid=case.grid,
                           npix=case.npix, covariates=cov, quantiles=[0.25, 0.5, 0.75], centi_int16=True)
    assert got["mean"].shape == (4, case.npix) and not np.isnan(got["mean"]).any()
    leaf = flat.split_val[(flat.left == 0) & (flat.right == 0)]
    assert got["mean"].min() >= leaf.min() and got["mean"].max() <= leaf.max()
    q = got["quantiles"]
    assert np.all(q[0] <= q[1]) and np.all(q[1] <= q[2])
    np.testing.assert_array_equal(got["iqr_i16"].ravel(), oracle.centi_int16(((q[2] - q[0]) / 2).ravel()))
    sel = np.sort(np.random.default_rng(1).choice(case.npix, 3000, replace=False))
    ref = oracle.predict_fused(flat.as_dict(), case.feature_names, loc_xy=loc[sel], obs_xy=case.obs_xy,
                               obs_z=case.obs_z, n_obs=case.n_obs, covariates={k: (v[sel] if v.ndim == 1 else v[:, sel])
                                                                              for k, v in cov.items()},
                               quantiles=[0.25, 0.5, 0.75], kdtree=True, threads=8)
    same_mean(got["mean"][:, sel], ref["mean"])
    np.testing.assert_array_equal(q[:, :, sel], ref["quantiles"])


def test_c5_scaleup_slice_device_vs_sampled_oracle():
    """10^4 stations, n = 20, F = 43, a 60-row band of the 10^4-column raster x 3 days with
    2 % of stations missing per day (cell-grid kNN + compact forest image)."""
    case = synth.make_case("c5", ndays=3)
    X, y = synth.training_table(case, lambda l, o, z, n: oracle.near_obs(l, o, z, n, kdtree=True, nthreads=8)[:2])
    flat = grow.grow_forest(X, y, case.feature_names, num_trees=60)
    model = rb.RangerForest(flat)
    assert model.info()["nodes"] > 1_000_000
    begin, npix = 4000 * case.ncol, 60 * case.ncol
    rng = np.random.default_rng(2)
    sel = np.sort(rng.choice(npix, 2500, replace=False))
    loc = case.locations(begin, npix)
    cov = {n: np.stack([case.cov_daily[n](loc, d) for d in range(3)]) for n in case.cov_names}
    got = rb.predict_fused(model, obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, grid=case.grid,
                           pix_begin=begin, npix=npix, covariates=cov)
    assert not np.isnan(got["mean"]).any()
    ref = oracle.predict_fused(flat.as_dict(), case.feature_names, loc_xy=loc[sel], obs_xy=case.obs_xy,
                               obs_z=case.obs_z, n_obs=case.n_obs, covariates={k: v[:, sel] for k, v in cov.items()},
                               kdtree=True, threads=8)
    same_mean(got["mean"][:, sel], ref["mean"])
