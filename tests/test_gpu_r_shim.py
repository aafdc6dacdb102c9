"""The R boundary on the GPU without R: every compute entry of rfsi_b200/R/src/r_shim.c is called the
way R's .Call would call it (tests/tools/mock_libR.c supplies SEXPs, PROTECT, R_alloc, Rf_error as a
longjmp to the top level) and its result is held to the oracle -- the same bars as the C-ABI tests:
indices / distances / observations / means / quantiles bit-equal."""
import numpy as np
import pytest

import oracle
from rfsi_b200 import synth
from rfsi_b200.api import feature_map
from tests.rmock import MockR, RError
from tests.util import same_mean

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    r = MockR()
    yield r
    r.gc()


@pytest.fixture(scope="module")
def case_model(R):
    case = synth.make_case("c1", ndays=3)
    case.obs_z[1, ::7] = np.nan
    X, y = synth.training_table(case, lambda l, o, z, n: oracle.near_obs(l, o, z, n)[:2])
    flat = synth.grow_forest_sklearn(X, y, case.feature_names, num_trees=40, quantreg=True, n_jobs=1)
    T = flat.ntrees
    n = np.diff(flat.node_offsets)
    rnv = np.full((int(n.max()), T), np.nan)          # ranger's random.node.values: max_nodes x num.trees
    for t in range(T):
        rnv[: n[t], t] = flat.random_node_values[flat.rnv_offsets[t]: flat.rnv_offsets[t] + n[t]]
    h = R.call("C_rfsi_forest_new", flat.node_offsets.astype(np.float64), flat.left, flat.right, flat.split_var,
               flat.split_val, rnv, len(case.feature_names))
    return case, flat, h


def test_near_obs_through_the_call_entry(R, case_model):
    case, _, _ = case_model
    loc = case.locations()
    res = R.call("C_rfsi_near_obs", loc, case.obs_xy, case.obs_z[0], case.n_obs, True, 0)
    assert list(res) == ["dist", "obs", "idx"]
    d, o, i, _ = oracle.near_obs(loc, case.obs_xy, case.obs_z[0], case.n_obs)
    np.testing.assert_array_equal(res["dist"], d)
    np.testing.assert_array_equal(res["obs"], o)
    np.testing.assert_array_equal(res["idx"], i)
    # the data.frame form near.obs() itself returns: dist1..distN, obs1..obsN, filled column by column
    df = R.call("C_rfsi_near_obs_df", loc, case.obs_xy, case.obs_z[0], case.n_obs, True, 0)
    n = case.n_obs
    assert list(df.columns) == [f"dist{j + 1}" for j in range(n)] + [f"obs{j + 1}" for j in range(n)]
    np.testing.assert_array_equal(df.to_numpy(), np.column_stack([d, o]))
    # too few stations: NA-filled columns and an R warning, not an error
    before = R.warning_count()
    res = R.call("C_rfsi_near_obs", loc[:50], case.obs_xy[:3], case.obs_z[0, :3], 5, True, 0)
    assert R.warning_count() == before + 1 and "fewer observations" in R.last_warning()
    assert np.isnan(res["dist"][:, 3:]).all() and (res["idx"][:, 3:] == np.iinfo(np.int32).min).all()
    # NA coordinate: R error carrying the engine's message; the context stays usable
    bad = loc[:200].copy(); bad[17, 1] = np.nan
    with pytest.raises(RError, match="NA coordinate"):
        R.call("C_rfsi_near_obs", bad, case.obs_xy, case.obs_z[0], case.n_obs, True, 0)
    res = R.call("C_rfsi_near_obs", loc[:200], case.obs_xy, case.obs_z[0], case.n_obs, True, 0)
    np.testing.assert_array_equal(res["dist"], d[:200])


def test_near_obs_days_through_the_call_entry(R, case_model):
    case, _, _ = case_model
    res = R.call("C_rfsi_near_obs_days", None, case.obs_xy, case.obs_z.T.copy(), case.n_obs, 0)
    S, n, D = case.obs_xy.shape[0], case.n_obs, case.obs_z.shape[0]
    dist = res[0].reshape(D, n, S)
    for day in range(D):
        ok = ~np.isnan(case.obs_z[day])
        d, _, _, _ = oracle.near_obs(case.obs_xy[ok], case.obs_xy[ok], case.obs_z[day][ok], n)
        np.testing.assert_array_equal(dist[day][:, ok].T, d)
        assert np.isnan(dist[day][:, ~ok]).all()


def test_predict_through_the_call_entry(R, case_model):
    case, flat, h = case_model
    loc = case.locations(0, 3000)
    d, o, _, _ = oracle.near_obs(loc, case.obs_xy, case.obs_z[0], case.n_obs)
    X = np.column_stack([d, o])
    fo = flat.as_dict()
    same_mean(R.call("C_rfsi_predict", h, X, None, 0), oracle.forest_predict(fo, X))
    probs = np.array([0.25, 0.5, 0.75])
    np.testing.assert_array_equal(R.call("C_rfsi_predict", h, X, probs, 0), oracle.forest_quantiles(fo, X, probs))
    Xn = X.copy(); Xn[5, 2] = np.nan                  # ranger: "Missing data in columns" is an error
    with pytest.raises(RError, match="missing data"):
        R.call("C_rfsi_predict", h, Xn, probs, 0)
    same_mean(R.call("C_rfsi_predict", h, X, None, 0), oracle.forest_predict(fo, X))


def test_pred_fused_and_pred_maps_through_the_call_entries(R, case_model):
    case, flat, h = case_model
    loc = case.locations()
    D = case.obs_z.shape[0]
    fmap = feature_map(case.feature_names, case.cov_names, case.n_obs)
    probs = np.array([0.25, 0.5, 0.75])
    res = R.call("C_rfsi_pred_fused", h, loc, case.obs_xy, case.obs_z.T.copy(), case.n_obs, [], fmap, probs, 0)
    ref = oracle.predict_fused(flat.as_dict(), case.feature_names, loc_xy=loc, obs_xy=case.obs_xy, obs_z=case.obs_z,
                               n_obs=case.n_obs, quantiles=list(probs))
    same_mean(res[0].T, ref["mean"])                                    # npix x ndays matrix
    q = res[1].reshape(len(probs), D, len(loc))
    np.testing.assert_array_equal(q, ref["quantiles"])
    maps = R.call("C_rfsi_pred_maps", h, loc, None, case.obs_xy, case.obs_z.T.copy(), case.n_obs, [], fmap, probs,
                  None, None, 0)
    def as_r_integer(x):                                # Int16 product -> R integer matrix, NAflag -> NA_integer_
        v = oracle.centi_int16(x).astype(np.int32).reshape(x.shape)
        return np.where(v == -32767, np.iinfo(np.int32).min, v)
    np.testing.assert_array_equal(maps[0].T, as_r_integer(ref["mean"]))
    np.testing.assert_array_equal(maps[1].T, as_r_integer((ref["quantiles"][2] - ref["quantiles"][0]) / 2))


def test_terminal_nodes_through_the_call_entry(R, case_model):
    case, flat, h = case_model
    loc = case.locations(0, 500)
    d, o, _, _ = oracle.near_obs(loc, case.obs_xy, case.obs_z[0], case.n_obs)
    X = np.column_stack([d, o])
    np.testing.assert_array_equal(R.call("C_rfsi_terminal_nodes", h, X, 0), oracle.forest_terminal_nodes(flat.as_dict(), X))
