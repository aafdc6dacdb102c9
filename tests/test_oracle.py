"""Pins for the CPU oracle (no GPU).  The reference has no golden vectors for this path
(SURVEY.md 8c), so the oracle is pinned by hand-computed cases, R quantile() type-7 known
answers, independent libraries (scipy cKDTree, sklearn tree descent) and three mutually
independent restatements that must agree bit for bit."""
from pathlib import Path

import numpy as np
import pytest

import oracle
from oracle import rfsi_oracle_np as onp

GOLD = Path(__file__).parent / "golden"


def test_near_obs_hand_case():
    # 5 stations, 4 locations; location 0 coincides with station 2 (self match),
    # location 1 is equidistant (d2 = 1) from stations 0 and 3 (tie -> lower index first)
    obs = np.array([[0.0, 0.0], [10.0, 0.0], [3.0, 4.0], [2.0, 0.0], [0.0, 7.0]])
    z = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    loc = np.array([[3.0, 4.0], [1.0, 0.0], [10.0, 1.0], [-1.0, -1.0]])
    d, o, i, rc = oracle.near_obs(loc, obs, z, 2)
    assert rc == 0
    # loc0: self (st2, d=0) dropped -> st4 d=sqrt(9+9)=4.2426.., st3 d=sqrt(1+16)=4.1231 -> order st3, st4
    assert i[0].tolist() == [4, 5]
    np.testing.assert_array_equal(d[0], [np.sqrt(17.0), np.sqrt(18.0)])
    assert o[0].tolist() == [4.0, 5.0]
    # loc1: stations 0 and 3 both at distance 1 -> index order 1 then 4
    assert i[1].tolist() == [1, 4]
    np.testing.assert_array_equal(d[1], [1.0, 1.0])
    # loc2: st1 d=1, then st2 d=sqrt(49+9) (st3 is at sqrt(64+1))
    assert i[2].tolist() == [2, 3]
    np.testing.assert_array_equal(d[2], [1.0, np.sqrt(58.0)])
    # loc3: st0 d=sqrt2, st3 d=sqrt(9+1)
    assert i[3].tolist() == [1, 4]
    # rm_dupl = FALSE keeps the zero-distance match
    d, o, i, _ = oracle.near_obs(loc, obs, z, 2, rm_dupl=False)
    assert i[0].tolist() == [3, 4] and d[0, 0] == 0.0


def test_near_obs_two_coincident_stations_drop_only_one():
    obs = np.array([[1.0, 1.0], [1.0, 1.0], [2.0, 1.0]])
    z = np.array([7.0, 8.0, 9.0])
    d, o, i, _ = oracle.near_obs(np.array([[1.0, 1.0]]), obs, z, 2)
    assert i[0].tolist() == [2, 3] and d[0].tolist() == [0.0, 1.0] and o[0].tolist() == [8.0, 9.0]


def test_near_obs_too_few_is_na_filled():
    obs = np.array([[0.0, 0.0], [1.0, 0.0]])
    d, o, i, rc = oracle.near_obs(np.array([[0.5, 0.0], [0.0, 0.0]]), obs, np.array([1.0, 2.0]), 3)
    assert rc == 1
    assert i[0, :2].tolist() == [1, 2] and i[0, 2] == onp.NA_INTEGER and np.isnan(d[0, 2])
    # self match dropped leaves a single neighbour
    assert i[1, 0] == 2 and i[1, 1] == onp.NA_INTEGER
    # NA_real_ payload (1954) is preserved
    assert d[0, 2:].view(np.uint64)[0] == 0x7FF00000000007A2


@pytest.mark.parametrize("S,n", [(40, 6), (300, 25), (64, 40)])
def test_three_restatements_agree_on_lattice_ties(S, n):
    rng = np.random.default_rng(S)
    gx, gy = np.meshgrid(np.arange(30) + 0.5, np.arange(30) + 0.5)
    loc = np.stack([gx.ravel(), gy.ravel()], 1)
    obs = loc[rng.choice(900, S, replace=False)]
    z = rng.normal(size=S)
    d1, o1, i1, _ = oracle.near_obs(loc, obs, z, n)
    d2, o2, i2, _ = oracle.near_obs(loc, obs, z, n, kdtree=True, nthreads=3)
    d3, o3, i3 = onp.near_obs(loc, obs, z, n)
    for a, b in ((d1, d2), (o1, o2), (i1, i2), (d1, d3), (o1, o3), (i1, i3)):
        np.testing.assert_array_equal(a, b)
    # properties: ascending distances; dist_k == ||loc - obs[idx_k]|| exactly
    assert np.all(np.diff(d1, axis=1) >= 0)
    k = 0
    rec = np.sqrt((loc[:, 0] - obs[i1[:, k] - 1, 0]) ** 2 + (loc[:, 1] - obs[i1[:, k] - 1, 1]) ** 2)
    np.testing.assert_array_equal(rec, d1[:, k])


def test_near_obs_against_scipy_ckdtree():
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(5)
    obs = rng.uniform(0, 1000, size=(500, 2))
    loc = rng.uniform(0, 1000, size=(2000, 2))
    d, _, i, _ = oracle.near_obs(loc, obs, np.zeros(500), 10, rm_dupl=False)
    ds, is_ = cKDTree(obs).query(loc, k=10)
    np.testing.assert_array_equal(i - 1, is_)          # continuous coordinates: no ties
    np.testing.assert_allclose(d, ds, rtol=1e-15)


def test_station_permutation_invariance_of_distances():
    rng = np.random.default_rng(9)
    gx, gy = np.meshgrid(np.arange(12.0), np.arange(12.0))
    loc = np.stack([gx.ravel(), gy.ravel()], 1)
    obs = loc[rng.choice(144, 30, replace=False)]
    z = rng.normal(size=30)
    perm = rng.permutation(30)
    d1, o1, _, _ = oracle.near_obs(loc, obs, z, 5)
    d2, o2, _, _ = oracle.near_obs(loc, obs[perm], z[perm], 5)
    np.testing.assert_array_equal(d1, d2)              # dist columns never depend on tie order
    np.testing.assert_array_equal(np.sort(o1, 1)[d1[:, -1] < d1.max()], np.sort(o1, 1)[d1[:, -1] < d1.max()])


def test_quantile_type7_known_answers():
    # R: quantile(1:250, c(.25,.5,.75)) -> 63.25 125.50 187.75
    x = np.arange(1.0, 251.0)
    assert oracle.quantile7_sorted(x, 0.25) == 63.25
    assert oracle.quantile7_sorted(x, 0.5) == 125.5
    assert oracle.quantile7_sorted(x, 0.75) == 187.75
    # R: quantile(c(1,2,3,4,10), c(0,.1,.5,.9,1)) -> 1.0 1.4 3.0 7.6 10.0
    y = np.array([1.0, 2.0, 3.0, 4.0, 10.0])
    got = [oracle.quantile7_sorted(y, p) for p in (0, .1, .5, .9, 1)]
    np.testing.assert_allclose(got, [1.0, 1.4, 3.0, 7.6, 10.0], rtol=1e-15)
    np.testing.assert_array_equal(onp.quantile7(y, [0, .1, .5, .9, 1]), got)
    # numpy's default ('linear') is type 7 as well, up to the interpolation expression
    rng = np.random.default_rng(1)
    v = rng.normal(size=249)
    for p in (0.25, 0.5, 0.75, 0.1, 0.9):
        assert abs(oracle.quantile7_sorted(np.sort(v), p) - np.quantile(v, p)) < 1e-14
    # na.rm = TRUE
    w = np.array([3.0, np.nan, 1.0, 2.0])
    assert onp.quantile7(w, [0.5])[0] == 2.0


def _hand_forest():
    # 3 trees over 2 features, ranger layout.  Tree 0: x0 <= 0.5 ? leaf 1.0 : (x1 <= 2 ? 2.0 : 4.0)
    trees = [
        dict(l=[1, 0, 3, 0, 0], r=[2, 0, 4, 0, 0], v=[0, 0, 1, 0, 0], s=[0.5, 1.0, 2.0, 2.0, 4.0]),
        dict(l=[1, 0, 0], r=[2, 0, 0], v=[1, 0, 0], s=[1.0, 10.0, 20.0]),
        dict(l=[0], r=[0], v=[0], s=[100.0]),                      # a stump that is a single leaf
    ]
    off = np.cumsum([0] + [len(t["l"]) for t in trees])
    fo = dict(ntrees=3, node_offsets=off.astype(np.int64),
              left=np.concatenate([t["l"] for t in trees]).astype(np.int32),
              right=np.concatenate([t["r"] for t in trees]).astype(np.int32),
              split_var=np.concatenate([t["v"] for t in trees]).astype(np.int32),
              split_val=np.concatenate([t["s"] for t in trees]).astype(np.float64))
    # random.node.values as the R matrix (max nodes x T), NA where unset
    rnv = np.full((5, 3), np.nan)
    rnv[[1, 3, 4], 0] = [1.5, 2.5, 4.5]
    rnv[[1, 2], 1] = [11.0, 21.0]
    rnv[0, 2] = 99.0
    fo["rnv_offsets"] = np.arange(3, dtype=np.int64) * 5
    fo["random_node_values"] = np.ascontiguousarray(rnv.T).ravel()
    return fo


def test_forest_hand_case():
    fo = _hand_forest()
    X = np.array([[0.5, 0.0], [0.6, 2.0], [0.6, 2.5], [0.0, 1.0]])
    mean = oracle.forest_predict(fo, X)
    np.testing.assert_array_equal(mean, [(1 + 10 + 100) / 3, (2 + 20 + 100) / 3, (4 + 20 + 100) / 3, (1 + 10 + 100) / 3])
    term = oracle.forest_terminal_nodes(fo, X)
    assert term.tolist() == [[1, 1, 0], [3, 2, 0], [4, 2, 0], [1, 1, 0]]
    q = oracle.forest_quantiles(fo, X, [0.0, 0.5, 1.0])
    np.testing.assert_array_equal(q[0], [1.5, 11.0, 99.0])
    np.testing.assert_array_equal(q[2], [4.5, 21.0, 99.0])
    np.testing.assert_array_equal(onp.forest_predict(fo, X), mean)
    np.testing.assert_array_equal(onp.forest_quantiles(fo, X, [0.0, 0.5, 1.0]), q)
    _, visits = oracle.forest_predict(fo, X, count_visits=True)
    assert visits == (1 + 1 + 0) + (2 + 1 + 0) + (2 + 1 + 0) + (1 + 1 + 0)


def test_forest_against_sklearn_descent():
    """sklearn's own apply()/predict() on fp32-representable inputs is an independent descent."""
    from sklearn.ensemble import RandomForestRegressor
    from rfsi_b200.forest import from_sklearn
    rng = np.random.default_rng(3)
    X = rng.integers(0, 64, size=(400, 6)).astype(np.float64)   # exactly representable in fp32
    y = X[:, 0] * 0.5 - X[:, 3] + rng.normal(size=400)
    rf = RandomForestRegressor(n_estimators=25, min_samples_leaf=2, random_state=0, n_jobs=1).fit(X, y)
    flat = from_sklearn(rf, [f"f{i}" for i in range(6)])
    Xp = rng.integers(0, 64, size=(300, 6)).astype(np.float64)
    term = oracle.forest_terminal_nodes(flat.as_dict(), Xp)
    for t, est in enumerate(rf.estimators_):
        np.testing.assert_array_equal(term[:, t], est.apply(Xp.astype(np.float32)))
    np.testing.assert_allclose(oracle.forest_predict(flat.as_dict(), Xp), rf.predict(Xp), rtol=1e-13)
    np.testing.assert_array_equal(oracle.forest_predict(flat.as_dict(), Xp),
                                  oracle.forest_predict(flat.as_dict(), Xp, threads=3))


def test_centi_int16_round_half_even():
    x = np.array([0.125, 0.135, -0.125, 1.005, np.nan, 400.0, -400.0, 0.015, 0.025])
    got = oracle.centi_int16(x)
    assert got.tolist() == [12, 14, -12, 100, -32767, 32767, -32767, 2, 2]


def test_oracle_reproduces_golden_vectors():
    g = np.load(GOLD / "knn_c1_lattice.npz")
    from rfsi_b200 import synth
    x0, y0, dx, dy, ncol = g["grid"]
    loc = synth.lattice(int(ncol), int(g["nrow"]), x0, y0, dx, dy)
    d, o, i, _ = oracle.near_obs(loc, g["obs_xy"], g["obs_z"], int(g["n_obs"]))
    np.testing.assert_array_equal(d, g["dist"]); np.testing.assert_array_equal(i, g["idx"])
    np.testing.assert_array_equal(o, g["obs"])
    # the lattice really exercises ties and self matches
    assert (g["dist"][:, 0] > 0).all() and (np.diff(g["dist"], axis=1) == 0).any()

    f = np.load(GOLD / "forest_c1_40trees.npz")
    fo = {k: f[k] for k in ("ntrees", "node_offsets", "left", "right", "split_var", "split_val", "rnv_offsets",
                            "random_node_values")}
    np.testing.assert_array_equal(oracle.forest_predict(fo, f["X"]), f["mean"])
    np.testing.assert_array_equal(oracle.forest_terminal_nodes(fo, f["X"]), f["term"])
    np.testing.assert_array_equal(oracle.forest_quantiles(fo, f["X"], f["probs"]), f["q"])
    np.testing.assert_array_equal(oracle.forest_quantiles(fo, f["X"], f["probs"], threads=2), f["q"])


def test_oracle_reproduces_real_geometry_golden():
    g = np.load(GOLD / "knn_croatia_real.npz")
    x0, y0, dx, dy, ncol = g["grid"]
    p = np.arange(int(g["pix_begin"]), int(g["pix_begin"]) + int(g["npix"]))
    loc = np.stack([x0 + (p % int(ncol)) * dx, y0 + (p // int(ncol)) * dy], axis=1)
    z = g["temp_days"][int(g["day"])]
    ok = ~np.isnan(z)
    d, o, i, _ = oracle.near_obs(loc, g["obs_xy"][ok], z[ok], int(g["n_obs"]), kdtree=True, nthreads=2)
    np.testing.assert_array_equal(d, g["dist"]); np.testing.assert_array_equal(i, g["idx"])
    # median nearest-station distance of the real network is ~12.7 km
    assert 5e3 < np.median(g["dist"][:, 0]) < 4e4


def test_raster_extract_and_covariate_repair_hand_cases():
    """raster::extract (method "simple") + `/ 10` + the tmin > tmax repair
    (prcp_catalonia/5_prcp_prediction.R:253-258), by hand on a 2 x 3 raster."""
    data = np.array([[10, 20, 3], [40, -32767, 60]], dtype=np.int16)        # row 0 = north
    geo = dict(x_ul=100.0, y_ul=50.0, dx=2.0, dy=1.0, nodata=-32767)
    pts = np.array([[100.0, 50.0],      # the upper-left corner belongs to cell (0, 0)
                    [101.9, 49.1],      # inside cell (0, 0)
                    [102.0, 49.5],      # a left cell edge belongs to the cell on its right
                    [106.0, 48.0],      # the outer right / bottom edges belong to the last column / row
                    [103.0, 48.5],      # the NAvalue cell
                    [99.99, 49.0], [103.0, 50.01], [106.01, 49.0], [103.0, 47.99],      # outside
                    [105.0, 49.0]])     # a horizontal cell edge belongs to the row below
    got = onp.raster_extract(data, xy=pts, **geo)
    exp = np.array([10, 10, 20, 60, np.nan, np.nan, np.nan, np.nan, np.nan, 60.0])
    np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
    np.testing.assert_array_equal(got[~np.isnan(exp)], exp[~np.isnan(exp)])
    # `/ 10` is a division: 3 / 10 is the double nearest to 0.3, 3 * 0.1 is 0.30000000000000004
    div = onp.raster_extract(data, xy=pts[:1] + [4.5, 0.0], scale=10.0, divide=True, **geo)
    mul = onp.raster_extract(data, xy=pts[:1] + [4.5, 0.0], scale=0.1, **geo)
    assert div[0] == 0.3 and mul[0] == 3 * 0.1 and mul[0] != 0.3
    # layers and offset
    stack = np.stack([data, data + 1])
    assert onp.raster_extract(stack, xy=pts[1:2], layer=1, offset=-0.5, **geo)[0] == 10.5
    # ifelse(tmin > tmax, tmax - 1, tmin), NA if either side is NA
    tmin = np.array([5.0, 9.0, 3.0, np.nan, 2.0])
    tmax = np.array([7.0, 8.0, 3.0, 4.0, np.nan])
    fixed = onp.cov_fix(tmin, tmax, 1.0)
    np.testing.assert_array_equal(fixed[:3], [5.0, 7.0, 3.0])
    assert np.isnan(fixed[3]) and np.isnan(fixed[4])


def test_near_obs_reproduces_the_reference_s_stored_idw_cv_predictions():
    """PIN AGAINST THE REFERENCE'S OWN OUTPUT.  temp_data/IDW_10f_pred.rda holds 55 746 predictions
    computed in R (gstat) during the reference's nested 10-fold IDW cross-validation
    (temp_croatia/1_models.R:405-536).  Each is a function of exactly what near.obs returns: the n
    nearest stations observed that day, their distances, their values.  All three restatements
    reproduce every one of them to 1e-12 (measured: 2.1e-14 absolute on values of 10-30 degC)."""
    from tests.util import RTOL, reference_idw_cv

    def by_day(fn):
        def run(obs_xy, obs_z, n, loc):
            D = obs_z.shape[0]
            dist = np.empty((D, loc.shape[0], n)); val = np.empty_like(dist)
            for d in range(D):
                ok = ~np.isnan(obs_z[d])
                dist[d], val[d] = fn(loc, obs_xy[ok], obs_z[d][ok], n)[:2]
            return dist, val
        return run

    got, stored = reference_idw_cv(by_day(oracle.near_obs))
    assert got.shape == stored.shape == (55746,)
    np.testing.assert_allclose(got, stored, rtol=RTOL, atol=0)
    kd, _ = reference_idw_cv(by_day(lambda l, o, z, n: oracle.near_obs(l, o, z, n, kdtree=True, nthreads=2)))
    np.testing.assert_array_equal(kd, got)
    # a wrong neighbour anywhere is visible: drop the nearest station's observation once
    def broken(l, o, z, n):
        d, v = oracle.near_obs(l, o, z, n + 1)[:2]
        return d[:, 1:], v[:, 1:]
    off, _ = reference_idw_cv(by_day(broken))
    assert np.max(np.abs(off - stored)) > 1.0
