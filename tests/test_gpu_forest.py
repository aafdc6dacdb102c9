"""Stage-2 parity on a real B200: forest / quantile-forest forward pass vs the CPU oracle.
Tolerance from north_star: 1e-12 relative in fp64 (the mean is summed in tree order, so we
also require bit equality; quantiles use R's exact interpolation expression)."""
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

import oracle
import rfsi_b200 as rb
from rfsi_b200 import synth
from rfsi_b200.forest import FlatForest
from tests.util import same_mean

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
RTOL = 1e-12


def _golden_forest():
    f = np.load(GOLD / "forest_c1_40trees.npz")
    flat = FlatForest(int(f["ntrees"]), f["node_offsets"], f["left"], f["right"], f["split_var"], f["split_val"],
                      [str(s) for s in f["feature_names"]], f["rnv_offsets"], f["random_node_values"])
    return f, flat


def test_golden_mean_terminal_nodes_quantiles():
    f, flat = _golden_forest()
    model = rb.RangerForest(flat)
    X = f["X"]
    mean = rb.predict(model, X).predictions
    same_mean(mean, f["mean"])
    np.testing.assert_array_equal(rb.predict(model, X, type="terminalNodes").predictions, f["term"])
    q = rb.predict(model, X, type="quantiles", quantiles=f["probs"]).predictions
    np.testing.assert_allclose(q, f["q"], rtol=RTOL, atol=0)
    np.testing.assert_array_equal(q, f["q"])


def test_predict_matches_columns_by_name_like_ranger():
    f, flat = _golden_forest()
    model = rb.RangerForest(flat)
    X = f["X"]
    df = pd.DataFrame({n: X[:, j] for j, n in enumerate(flat.feature_names)})
    df = df[df.columns[::-1]].assign(sim1=0.0, fold=True)               # shuffled + extra columns
    same_mean(rb.predict(model, df).predictions, f["mean"])
    assert rb.predict(model, df, type="quantiles").predictions.shape == (X.shape[0], 3)   # default 0.1/0.5/0.9


def test_missing_data_is_an_error_like_ranger():
    f, flat = _golden_forest()
    model = rb.RangerForest(flat)
    X = f["X"].copy()
    X[5, 3] = np.nan
    with pytest.raises(rb.RfsiError) as e:
        rb.predict(model, X)
    assert e.value.code == -4
    plain = FlatForest(flat.ntrees, flat.node_offsets, flat.left, flat.right, flat.split_var, flat.split_val,
                       flat.feature_names)
    with pytest.raises(ValueError):
        rb.predict(rb.RangerForest(plain), f["X"], type="quantiles")


@pytest.mark.parametrize("image", ["3", "2", "1", "0"])
def test_missing_data_in_quantiles_and_terminal_nodes(image, monkeypatch):
    """NA in a used column with type="quantiles" / "terminalNodes" (ranger: "Missing data in columns"): the C ABI
    returns RFSI_E_NA_INPUT, the NA rows come back NA, every other row is still right, and the device context
    stays usable (the rows' leaf ids are never written: the quantile pass must skip them, not gather from them)."""
    import ctypes as C
    from rfsi_b200 import _lib
    monkeypatch.setenv("RFSI_B200_COMPACT", image)
    f, flat = _golden_forest()
    model = rb.RangerForest(flat)
    lib = _lib.load()
    X = f["X"].copy()
    bad = np.array([0, 5, 77, len(X) - 1])
    X[bad, [3, 0, 1, 2]] = np.nan
    Xt = np.ascontiguousarray(X.T)
    npix = len(X)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    good = np.setdiff1d(np.arange(npix), bad)
    probs = np.ascontiguousarray(f["probs"], dtype=np.float64)
    for _ in range(2):                               # twice: the first failure must not poison the context
        q = np.full((len(probs), npix), 7.0)
        rc = lib.rfsi_forest_quantiles(model.handle, ptr(Xt), npix, npix, ptr(probs), len(probs), ptr(q), 0)
        assert rc == -4, _lib.last_error()
        assert np.isnan(q[:, bad]).all()
        np.testing.assert_array_equal(q[:, good].T, f["q"][good])
        term = np.full((flat.ntrees, npix), 7, dtype=np.int32)
        assert lib.rfsi_forest_terminal_nodes(model.handle, ptr(Xt), npix, npix, ptr(term), 0) == -4
        assert (term[:, bad] == np.iinfo(np.int32).min).all()            # NA_integer_
        np.testing.assert_array_equal(term[:, good].T, f["term"][good])
        mean = np.full(npix, 7.0)
        assert lib.rfsi_forest_predict(model.handle, ptr(Xt), npix, npix, ptr(mean), 0) == -4
        assert np.isnan(mean[bad]).all()
        same_mean(mean[good], f["mean"][good])
    # and a clean call on the same context afterwards
    np.testing.assert_array_equal(rb.predict(model, f["X"], type="quantiles", quantiles=f["probs"]).predictions, f["q"])
    np.testing.assert_array_equal(rb.predict(model, f["X"], type="terminalNodes").predictions, f["term"])


def _random_forest_case(seed, n_train, F, T, min_leaf, quantreg=True):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n_train, F))
    y = X[:, 0] * 2 + np.sin(X[:, 1 % F] * 3) + rng.normal(scale=0.3, size=n_train)
    flat = synth.grow_forest_sklearn(X, y, [f"f{i}" for i in range(F)], num_trees=T, min_node_size=min_leaf,
                                     quantreg=quantreg, n_jobs=8, seed=seed)
    return flat, rng


@pytest.mark.parametrize("F,T,npix", [(2, 1, 100), (12, 7, 1000), (50, 33, 5000), (80, 250, 3000), (130, 9, 700)])
def test_parity_many_shapes(F, T, npix):
    """T not a multiple of the trees-in-flight count, single tree, F up to the 128-pixel tile path."""
    flat, rng = _random_forest_case(F * 1000 + T, 600, F, T, 3)
    model = rb.RangerForest(flat)
    X = rng.normal(size=(npix, F))
    fo = flat.as_dict()
    same_mean(rb.predict(model, X).predictions, oracle.forest_predict(fo, X, threads=8))
    np.testing.assert_array_equal(rb.predict(model, X, type="terminalNodes").predictions,
                                  oracle.forest_terminal_nodes(fo, X))
    probs = [0.0, 0.1, 0.25, 0.5, 0.75, 0.9, 1.0]
    np.testing.assert_array_equal(rb.predict(model, X, type="quantiles", quantiles=probs).predictions,
                                  oracle.forest_quantiles(fo, X, probs, threads=8))


def test_quantiles_with_na_samples_and_duplicates():
    """na.rm=TRUE path: terminal nodes whose random.node.values cell was never set."""
    flat, rng = _random_forest_case(77, 300, 6, 64, 5)
    rnv = flat.random_node_values.copy()
    leaves = np.flatnonzero((flat.left == 0) & (flat.right == 0))
    # knock out a third of the leaf samples, duplicate values elsewhere
    for t in range(flat.ntrees):
        a, b = flat.node_offsets[t], flat.node_offsets[t + 1]
        lt = leaves[(leaves >= a) & (leaves < b)] - a
        rnv[flat.rnv_offsets[t] + lt[::3]] = np.nan
        rnv[flat.rnv_offsets[t] + lt[1::3]] = 1.25
    flat.random_node_values = rnv
    model = rb.RangerForest(flat)
    X = rng.normal(size=(2000, 6))
    probs = [0.25, 0.5, 0.75]
    got = rb.predict(model, X, type="quantiles", quantiles=probs).predictions
    np.testing.assert_array_equal(got, oracle.forest_quantiles(flat.as_dict(), X, probs, threads=8))


@pytest.mark.parametrize("T,npix", [(100, 700), (300, 500), (512, 300), (600, 300), (1024, 130)])
def test_quantiles_every_samples_per_lane_variant(T, npix):
    """The quantile kernel is instantiated per ceil(T/32) in {1,2,4,8,16,32}: cover 4, 16 and 32
    (1024 trees is the documented maximum), with row counts that leave warps and CTAs idle."""
    flat, rng = _random_forest_case(9000 + T, 200, 4, T, 4)
    model = rb.RangerForest(flat)
    X = rng.normal(size=(npix, 4))
    probs = [0.0, 0.05, 0.25, 0.5, 0.75, 0.99, 1.0]
    got = rb.predict(model, X, type="quantiles", quantiles=probs).predictions
    np.testing.assert_array_equal(got, oracle.forest_quantiles(flat.as_dict(), X, probs, threads=8))
    same_mean(rb.predict(model, X).predictions, oracle.forest_predict(flat.as_dict(), X, threads=8))


def test_quantiles_more_than_1024_trees_is_an_error():
    flat, rng = _random_forest_case(1, 60, 2, 1025, 10)
    model = rb.RangerForest(flat)
    with pytest.raises(rb.RfsiError) as e:
        rb.predict(model, rng.normal(size=(10, 2)), type="quantiles", quantiles=[0.5])
    assert e.value.code == -3 and "1024" in str(e.value)
    assert rb.predict(model, rng.normal(size=(10, 2))).predictions.shape == (10,)   # the mean still works


@pytest.mark.parametrize("mode", ["coarse", "two_values", "one_outlier", "all_equal", "zero_inflated", "zero_inflated_close", "top_heavy"])
def test_quantiles_heavy_ties_rebucketing(mode):
    """More than 32 equal / near-equal samples in one histogram bucket: the re-bucketing loop
    and its all-equal exit (qrf.cuh) against the oracle's sort."""
    flat, rng = _random_forest_case(4242, 400, 5, 300, 3)
    rnv = flat.random_node_values.copy()
    ok = ~np.isnan(rnv)
    if mode == "coarse":
        rnv[ok] = np.round(rnv[ok], 0)                       # a handful of distinct values
    elif mode == "two_values":
        rnv[ok] = np.where(rnv[ok] > 0, 1.0, 1.0 + 2.0 ** -40)   # one bucket at level 0, split at level 1
    elif mode == "one_outlier":
        rnv[ok] = 3.5 + 1e-9 * rng.integers(0, 3, size=int(ok.sum()))   # three values 1e-9 apart ...
        rnv[flat.rnv_offsets[0]:flat.rnv_offsets[1]] = 1.0e6             # ... and tree 0 far away
    elif mode == "zero_inflated":                            # daily precipitation: mostly exact zeros, a continuous tail
        rnv[ok] = np.where(rng.random(int(ok.sum())) < 0.6, 0.0, np.abs(rnv[ok]))
    elif mode == "zero_inflated_close":                      # ... whose smallest values share the zeros' bucket
        u = rng.random(int(ok.sum()))
        rnv[ok] = np.where(u < 0.5, 0.0, np.where(u < 0.6, 1e-7 * rng.integers(1, 40, size=int(ok.sum())), 50.0 * np.abs(rnv[ok])))
    elif mode == "top_heavy":                                # the heavy value is the window's largest: falls back to re-bucketing
        u = rng.random(int(ok.sum()))
        rnv[ok] = np.where(u < 0.5, 9.0, np.where(u < 0.6, 9.0 - 1e-7 * rng.integers(1, 40, size=int(ok.sum())), -50.0 * np.abs(rnv[ok])))
    else:
        rnv[ok] = -7.25
    flat.random_node_values = rnv
    model = rb.RangerForest(flat)
    X = rng.normal(size=(1500, 5))
    probs = [0.0, 0.1, 0.25, 0.5, 0.75, 0.9, 1.0]
    got = rb.predict(model, X, type="quantiles", quantiles=probs).predictions
    np.testing.assert_array_equal(got, oracle.forest_quantiles(flat.as_dict(), X, probs, threads=8))


def test_deep_skinny_tree_and_chunked_rows(monkeypatch):
    """A degenerate chain tree (depth 300) and many row chunks through the host API (several staging threads,
    each with its own stream: engine.cu "host staging")."""
    monkeypatch.setenv("RFSI_B200_PREDICT_CHUNK", "65536")
    depth = 300
    left = np.zeros(2 * depth + 1, dtype=np.int32); right = left.copy()
    var = left.copy(); val = np.zeros(2 * depth + 1)
    for k in range(depth):                       # node 2k: internal, children 2k+1 (leaf) and 2k+2
        left[2 * k], right[2 * k], val[2 * k] = 2 * k + 1, 2 * k + 2, float(k)
        val[2 * k + 1] = 1000.0 + k
    val[2 * depth] = -1.0
    flat = FlatForest(1, np.array([0, 2 * depth + 1], dtype=np.int64), left, right, var, val, ["x"])
    model = rb.RangerForest(flat)
    assert model.info()["max_depth"] == depth
    rng = np.random.default_rng(0)
    X = rng.uniform(-5, depth + 5, size=(900_000, 1))
    got = rb.predict(model, X).predictions
    exp = np.where(X[:, 0] > depth - 1, -1.0, 1000.0 + np.maximum(np.ceil(X[:, 0]), 0))
    np.testing.assert_array_equal(got, exp)


def test_forest_properties_at_simulation_size():
    """simulation.R scripted size: 250 000 rows x 250 trees x F=50; checked against the
    threaded oracle on a sample and through invariants everywhere."""
    flat, rng = _random_forest_case(5, 2000, 50, 250, 5)
    model = rb.RangerForest(flat)
    X = rng.normal(size=(250_000, 50))
    mean = rb.predict(model, X).predictions
    q = rb.predict(model, X, type="quantiles", quantiles=[0.25, 0.5, 0.75]).predictions
    leaf_vals = flat.split_val[(flat.left == 0) & (flat.right == 0)]
    assert mean.min() >= leaf_vals.min() and mean.max() <= leaf_vals.max()
    assert np.all(np.diff(q, axis=1) >= 0)                       # quantiles monotone in p
    sel = rng.choice(250_000, 5000, replace=False)
    same_mean(mean[sel], oracle.forest_predict(flat.as_dict(), X[sel], threads=8))
    np.testing.assert_array_equal(q[sel], oracle.forest_quantiles(flat.as_dict(), X[sel], [0.25, 0.5, 0.75], threads=8))
    # idempotence / determinism
    np.testing.assert_array_equal(mean, rb.predict(model, X).predictions)


def test_wide_and_compact_node_formats_agree(monkeypatch):
    """The sector image (default), the blocked 4-byte image, the 8-byte rank-quantised image and the
    16-byte fp64 image make identical decisions."""
    flat, rng = _random_forest_case(11, 3000, 20, 64, 2)
    X = rng.normal(size=(20000, 20))
    X[::7, 3] = flat.split_val[flat.left != 0][:len(X[::7])][:X[::7].shape[0]]   # features exactly ON thresholds
    exp = oracle.forest_predict(flat.as_dict(), X, threads=8)
    probs = [0.25, 0.5, 0.75]
    expq = oracle.forest_quantiles(flat.as_dict(), X, probs, threads=8)
    for compact in ("3", "2", "1", "0"):
        monkeypatch.setenv("RFSI_B200_COMPACT", compact)
        model = rb.RangerForest(flat)
        assert model.info()["nodes"] > 0 and model.info()["image"] == int(compact)
        same_mean(rb.predict(model, X).predictions, exp)
        np.testing.assert_array_equal(rb.predict(model, X, type="quantiles", quantiles=probs).predictions, expq)
        np.testing.assert_array_equal(rb.predict(model, X, type="terminalNodes").predictions,
                                      oracle.forest_terminal_nodes(flat.as_dict(), X))
