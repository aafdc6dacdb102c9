"""Stage-1 parity on a real B200: CUDA near.obs (through the C ABI) vs the CPU oracle.
Neighbour indices and distances must be bit-exact (north_star: indices bit-exact,
distances within 1e-12 relative -- we require equality)."""
from pathlib import Path

import numpy as np
import pytest

import oracle
import rfsi_b200 as rb
from rfsi_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def _gpu(loc, obs, z, n, **kw):
    df, idx = rb.near_obs(loc, np.column_stack([obs, z]), zcol=2, n_obs=n, return_idx=True, **kw)
    d = np.stack([df[f"dist{j + 1}"].to_numpy() for j in range(n)], 1)
    o = np.stack([df[f"obs{j + 1}"].to_numpy() for j in range(n)], 1)
    assert list(df.columns) == [f"dist{j + 1}" for j in range(n)] + [f"obs{j + 1}" for j in range(n)]
    return d, o, idx


def _same(a, b):
    np.testing.assert_array_equal(np.asarray(a).view(np.uint64) if a.dtype == np.float64 else a,
                                  np.asarray(b).view(np.uint64) if b.dtype == np.float64 else b)


def test_hand_case():
    obs = np.array([[0.0, 0.0], [10.0, 0.0], [3.0, 4.0], [2.0, 0.0], [0.0, 7.0]])
    z = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    loc = np.array([[3.0, 4.0], [1.0, 0.0], [10.0, 1.0], [-1.0, -1.0]])
    d, o, i = _gpu(loc, obs, z, 2)
    assert i.tolist() == [[4, 5], [1, 4], [2, 3], [1, 4]]
    ed, eo, ei, _ = oracle.near_obs(loc, obs, z, 2)
    _same(d, ed); _same(o, eo); _same(i, ei)
    d, o, i = _gpu(loc, obs, z, 2, rm_dupl=False)
    assert i[0].tolist() == [3, 4] and d[0, 0] == 0.0


def test_golden_lattice_ties_and_self_matches():
    g = np.load(GOLD / "knn_c1_lattice.npz")
    x0, y0, dx, dy, ncol = g["grid"]
    loc = synth.lattice(int(ncol), int(g["nrow"]), x0, y0, dx, dy)
    d, o, i = _gpu(loc, g["obs_xy"], g["obs_z"], int(g["n_obs"]))
    _same(d, g["dist"]); _same(i, g["idx"]); _same(o, g["obs"])


def test_golden_utm_geometry():
    g = np.load(GOLD / "knn_c3_utm.npz")
    case = synth.make_case("c3", ndays=1)
    loc = case.locations(0, int(g["npix"]))
    d, o, i = _gpu(loc, g["obs_xy"], g["obs_z"], int(g["n_obs"]))
    _same(d, g["dist"]); _same(i, g["idx"]); _same(o, g["obs"])


@pytest.mark.parametrize("S,n,seed", [(500, 40, 1), (3000, 25, 2), (1024, 10, 3), (1025, 1, 4), (7, 6, 5)])
def test_lattice_parity_many_shapes(S, n, seed):
    """sim_500.R shape (n = 40), multi-tile station sets (S > 1024), tile edges, n.obs+1 == S."""
    rng = np.random.default_rng(seed)
    side = 120
    loc = synth.lattice(side, side)
    obs = loc[rng.choice(side * side, S, replace=False)]
    z = rng.normal(size=S)
    d, o, i = _gpu(loc, obs, z, n)
    ed, eo, ei, _ = oracle.near_obs(loc, obs, z, n, kdtree=True, nthreads=8)
    _same(d, ed); _same(i, ei); _same(o, eo)


def test_too_few_observations_are_na_filled_with_warning():
    obs = np.array([[0.0, 0.0], [1.0, 0.0]])
    z = np.array([1.0, 2.0])
    loc = np.array([[0.5, 0.0], [0.0, 0.0]])
    with pytest.warns(UserWarning):
        d, o, i = _gpu(loc, obs, z, 3)
    ed, eo, ei, rc = oracle.near_obs(loc, obs, z, 3)
    assert rc == 1
    _same(d, ed); _same(i, ei); _same(o, eo)
    assert d.view(np.uint64)[0, 2] == 0x7FF00000000007A2 and i[0, 2] == np.iinfo(np.int32).min


def test_na_observation_values_propagate_and_na_coordinates_are_refused():
    obs = np.array([[0.0, 0.0], [1.0, 0.0], [5.0, 5.0]])
    z = np.array([1.0, np.nan, 3.0])
    loc = np.array([[0.9, 0.0]])
    d, o, i = _gpu(loc, obs, z, 2)
    assert i[0].tolist() == [2, 1] and np.isnan(o[0, 0]) and o[0, 1] == 1.0
    with pytest.raises(rb.RfsiError) as e:
        _gpu(np.array([[np.nan, 0.0]]), obs, z, 1)
    assert e.value.code == -4
    with pytest.raises(rb.RfsiError):
        _gpu(loc, np.array([[0.0, np.nan], [1.0, 1.0]]), np.array([1.0, 2.0]), 1)


def test_empty_and_ragged_inputs():
    obs = np.array([[0.0, 0.0], [1.0, 0.0], [5.0, 5.0]])
    z = np.array([1.0, 2.0, 3.0])
    d, o, i = _gpu(np.zeros((0, 2)), obs, z, 2)
    assert d.shape == (0, 2)
    # not a multiple of the CTA width, continuous coordinates
    rng = np.random.default_rng(0)
    loc = rng.uniform(-10, 10, size=(1000 + 37, 2))
    obs = rng.uniform(-10, 10, size=(333, 2))
    z = rng.normal(size=333)
    d, o, i = _gpu(loc, obs, z, 12)
    ed, eo, ei, _ = oracle.near_obs(loc, obs, z, 12)
    _same(d, ed); _same(i, ei); _same(o, eo)


def test_training_shape_locations_equal_observations():
    """temp_croatia/1_models.R:1040-1045: every station's own observation is excluded."""
    rng = np.random.default_rng(11)
    xy = rng.uniform(0, 1e5, size=(150, 2))
    z = rng.normal(size=150)
    d, o, i = _gpu(xy, xy, z, 10)
    assert (d[:, 0] > 0).all() and (i != np.arange(1, 151)[:, None]).all()
    ed, eo, ei, _ = oracle.near_obs(xy, xy, z, 10)
    _same(d, ed); _same(i, ei)


def test_chunked_large_call_matches_properties():
    """> 2^20 locations (two chunks, double-buffered streams): checked through
    size-independent properties + a kd-tree oracle sample."""
    rng = np.random.default_rng(21)
    side = 1100                                   # 1.21 M cells
    loc = synth.lattice(side, side)
    obs = loc[rng.choice(side * side, 2000, replace=False)]
    z = rng.normal(size=2000)
    d, o, i = _gpu(loc, obs, z, 8)
    assert np.all(np.diff(d, axis=1) >= 0) and (d[:, 0] > 0).all()
    rec = np.sqrt((loc[:, 0] - obs[i[:, 3] - 1, 0]) ** 2 + (loc[:, 1] - obs[i[:, 3] - 1, 1]) ** 2)
    np.testing.assert_array_equal(rec, d[:, 3])
    np.testing.assert_array_equal(z[i - 1], o)
    sel = rng.choice(side * side, 20000, replace=False)
    ed, eo, ei, _ = oracle.near_obs(loc[sel], obs, z, 8, kdtree=True, nthreads=8)
    _same(d[sel], ed); _same(i[sel], ei)


@pytest.mark.parametrize("case", ["shuffled", "outside", "collinear", "clustered", "forced_small"])
def test_cell_grid_search_is_exact_in_awkward_geometries(case, monkeypatch):
    """The pruned search (default for S >= 128) must be the same function of the data as the full scan:
    incoherent location order, locations far outside the station bbox, degenerate bboxes,
    strong clustering, and the cell path forced onto a small lattice with ties."""
    rng = np.random.default_rng(hash(case) % 1000)
    S, n = 4000, 15
    obs = rng.uniform(0, 5000, size=(S, 2))
    loc = rng.uniform(0, 5000, size=(30000, 2))
    if case == "outside":
        loc = rng.uniform(-20000, 30000, size=(30000, 2))
    elif case == "collinear":
        obs[:, 1] = 17.0
        obs[:100, 0] = 3.0                         # 100 coincident stations
    elif case == "clustered":
        obs[: S // 2] = rng.normal(2500, 5, size=(S // 2, 2))
        loc[:15000] = rng.normal(2500, 8, size=(15000, 2))
    elif case == "forced_small":
        monkeypatch.setenv("RFSI_B200_CELLS_MIN", "1")
        side = 90
        loc = synth.lattice(side, side)
        obs = loc[rng.choice(side * side, 700, replace=False)]
        S = 700
    z = rng.normal(size=S)
    d, o, i = _gpu(loc, obs, z, n)
    ed, eo, ei, _ = oracle.near_obs(loc, obs, z, n, kdtree=True, nthreads=8)
    _same(d, ed); _same(i, ei); _same(o, eo)


def test_golden_real_croatia_station_geometry():
    """Real inputs (stfdf.rda stations, dem.tif grid geometry, day 5), oracle outputs."""
    g = np.load(GOLD / "knn_croatia_real.npz")
    x0, y0, dx, dy, ncol = g["grid"]
    p = np.arange(int(g["pix_begin"]), int(g["pix_begin"]) + int(g["npix"]))
    loc = np.stack([x0 + (p % int(ncol)) * dx, y0 + (p // int(ncol)) * dy], axis=1)
    z = g["temp_days"][int(g["day"])]
    ok = ~np.isnan(z)
    assert 140 <= ok.sum() <= 157
    d, o, i = _gpu(loc, g["obs_xy"][ok], z[ok], int(g["n_obs"]))
    _same(d, g["dist"]); _same(i, g["idx"]); _same(o, g["obs"])


def test_near_obs_days_training_table_and_prediction_shapes():
    """All days in one call == the scripts' per-day loop (1_models.R:1034-1047): training
    shape (locations are the stations, own observation excluded, silent stations NA) and
    prediction shape (explicit locations), with enough missing stations on one day to force
    the in-kernel full-scan fallback."""
    rng = np.random.default_rng(8)
    S, D, n = 157, 6, 10
    xy = rng.uniform(0, 4e5, size=(S, 2))
    z = rng.normal(size=(D, S))
    for d in range(D):
        z[d, rng.choice(S, rng.integers(3, 9), replace=False)] = np.nan
    z[3, rng.choice(S, 120, replace=False)] = np.nan            # 37 stations left on day 3
    dist, obs, idx = rb.near_obs_days(xy, z, n, return_idx=True)
    assert dist.shape == (D, S, n)
    loc = rng.uniform(0, 4e5, size=(5000, 2))
    pd_, po_ = rb.near_obs_days(xy, z, n, locations=loc)
    for d in range(D):
        ok = ~np.isnan(z[d])
        ed, eo, ei, _ = oracle.near_obs(xy[ok], xy[ok], z[d][ok], n)
        _same(dist[d][ok], ed); _same(obs[d][ok], eo)
        np.testing.assert_array_equal(idx[d][ok], np.flatnonzero(ok)[ei - 1] + 1)   # original station numbering
        assert np.isnan(dist[d][~ok]).all() and (idx[d][~ok] == np.iinfo(np.int32).min).all()
        assert (dist[d][ok][:, 0] > 0).all()
        ed, eo, _, _ = oracle.near_obs(loc, xy[ok], z[d][ok], n)
        _same(pd_[d], ed); _same(po_[d], eo)


def test_cuda_near_obs_reproduces_the_reference_s_stored_idw_cv_predictions():
    """The CUDA kNN against an output vector the REFERENCE produced: the 55 746 stored predictions of
    its nested 10-fold IDW cross-validation (temp_croatia/1_models.R:405-536, temp_data/IDW_10f_pred.rda,
    gstat in R) are functions of exactly near.obs's output -- see tests/util.py::reference_idw_cv."""
    from tests.util import RTOL, reference_idw_cv
    got, stored = reference_idw_cv(lambda obs_xy, obs_z, n, loc: rb.near_obs_days(obs_xy, obs_z, n, locations=loc))
    assert got.shape == stored.shape == (55746,)
    np.testing.assert_allclose(got, stored, rtol=RTOL, atol=0)


def test_near_obs_f32_twin():
    """float buffers in/out: same neighbours as the fp64 call on the same coordinates,
    distances/observations within 1e-5 relative (north_star's fp32 mode)."""
    import ctypes as C
    from rfsi_b200 import _lib
    rng = np.random.default_rng(4)
    loc = rng.uniform(0, 1000, size=(5000, 2)).astype(np.float32)
    obs = rng.uniform(0, 1000, size=(300, 2)).astype(np.float32)
    z = rng.normal(size=300).astype(np.float32)
    n = 9
    dist = np.empty((n, 5000), dtype=np.float32); val = np.empty((n, 5000), dtype=np.float32)
    idx = np.empty((n, 5000), dtype=np.int32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    lx, ly = np.ascontiguousarray(loc[:, 0]), np.ascontiguousarray(loc[:, 1])
    ox, oy = np.ascontiguousarray(obs[:, 0]), np.ascontiguousarray(obs[:, 1])
    rc = _lib.load().rfsi_near_obs_f32(P(lx), P(ly), 5000, P(ox), P(oy), P(z), 300, n, 1, P(dist), P(val), P(idx), 0)
    assert rc == 0
    ed, eo, ei, _ = oracle.near_obs(loc.astype(np.float64), obs.astype(np.float64), z.astype(np.float64), n)
    np.testing.assert_array_equal(idx.T, ei)
    np.testing.assert_allclose(dist.T, ed, rtol=1e-5)
    np.testing.assert_array_equal(val.T, eo.astype(np.float32))


@pytest.mark.parametrize("case", ["lattice_ties", "utm_1km", "clustered", "few_stations_brute", "n40"])
def test_fp32_mode_neighbours_are_bit_exact(case):
    """fp32 mode (rfsi_near_obs_f32 / precision="fp32"): fp32 distances only PREFILTER the cell-grid search, every
    survivor is re-ranked by its exact fp64 distance, so the neighbour indices -- tie order included -- are those of
    the fp64 search on the same float-representable coordinates; distances within north_star's 1e-5 (in fact they are
    the fp64 distances rounded to float)."""
    rng = np.random.default_rng(11)
    n = 8
    if case == "lattice_ties":                   # every distance class is shared by several stations
        g = np.stack(np.meshgrid(np.arange(40.0), np.arange(40.0)), -1).reshape(-1, 2)
        obs = g[rng.choice(len(g), 700, replace=False)]
        loc = g
    elif case == "utm_1km":                      # UTM-sized coordinates: 24-bit floats resolve 0.5 m at 5e6
        obs = np.column_stack([rng.uniform(4.0e5, 8.0e5, 900), rng.uniform(4.7e6, 5.2e6, 900)])
        loc = np.column_stack([rng.uniform(4.0e5, 8.0e5, 20000), rng.uniform(4.7e6, 5.2e6, 20000)])
    elif case == "clustered":                    # dense clusters: many candidates within one float ulp of the k-th best
        c = rng.uniform(0, 1000, size=(12, 2))
        obs = (c[rng.integers(0, 12, 1500)] + rng.normal(scale=0.01, size=(1500, 2)))
        loc = (c[rng.integers(0, 12, 6000)] + rng.normal(scale=0.02, size=(6000, 2)))
    elif case == "few_stations_brute":           # below the cell-grid threshold: the brute-force kernel, float I/O only
        obs = rng.uniform(0, 100, size=(87, 2)); loc = rng.uniform(0, 100, size=(5000, 2))
    else:
        n = 40
        obs = rng.uniform(0, 200, size=(500, 2)); loc = synth.lattice(200, 200).astype(np.float64)
    obs = obs.astype(np.float32).astype(np.float64); loc = loc.astype(np.float32).astype(np.float64)
    z = rng.normal(size=len(obs)).astype(np.float32).astype(np.float64)
    df, idx = rb.near_obs(loc, np.column_stack([obs, z]), zcol=2, n_obs=n, return_idx=True, precision="fp32")
    ed, eo, ei, _ = oracle.near_obs(loc, obs, z, n)
    np.testing.assert_array_equal(idx, ei)
    got_d = np.column_stack([df[f"dist{j + 1}"] for j in range(n)]); got_o = np.column_stack([df[f"obs{j + 1}"] for j in range(n)])
    assert got_d.dtype == np.float32
    np.testing.assert_array_equal(got_d, ed.astype(np.float32))        # the fp64 distance, rounded once
    np.testing.assert_allclose(got_d, ed, rtol=1e-5)                    # north_star's fp32 tolerance
    np.testing.assert_array_equal(got_o, eo.astype(np.float32))


def test_fp32_mode_na_and_too_few():
    """The fp32 entry keeps the fp64 entry's contract: NA coordinates are RFSI_E_NA_INPUT, short lists are NA-filled
    with a warning, chunked calls equal one call."""
    import ctypes as C
    from rfsi_b200 import _lib
    rng = np.random.default_rng(5)
    obs = rng.uniform(0, 10, size=(5, 2)).astype(np.float32); z = rng.normal(size=5).astype(np.float32)
    loc = rng.uniform(0, 10, size=(300, 2)).astype(np.float32)
    with pytest.warns(UserWarning):
        df = rb.near_obs(loc, np.column_stack([obs, z]), zcol=2, n_obs=7, precision="fp32")
    assert np.isnan(df["dist6"]).all() and np.isnan(df["obs7"]).all() and not np.isnan(df["dist4"]).any()
    bad = loc.copy(); bad[17, 0] = np.nan
    with pytest.raises(rb.RfsiError) as e:
        rb.near_obs(bad, np.column_stack([obs, z]), zcol=2, n_obs=2, precision="fp32")
    assert e.value.code == _lib.RFSI_E_NA_INPUT
    # many chunks, several staging threads
    obs = rng.uniform(0, 1000, size=(400, 2)).astype(np.float32); z = rng.normal(size=400).astype(np.float32)
    loc = rng.uniform(0, 1000, size=(150_000, 2)).astype(np.float32)
    df, idx = rb.near_obs(loc, np.column_stack([obs, z]), zcol=2, n_obs=5, return_idx=True, precision="fp32")
    ed, eo, ei, _ = oracle.near_obs(loc.astype(np.float64), obs.astype(np.float64), z.astype(np.float64), 5, kdtree=True, nthreads=8)
    np.testing.assert_array_equal(idx, ei)
    np.testing.assert_array_equal(np.column_stack([df[f"dist{j + 1}"] for j in range(5)]), ed.astype(np.float32))
