"""Property tests (hypothesis) that keep the three CPU restatements in lock step on adversarial
small inputs: integer lattices (ties everywhere), coincident stations, short station lists."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle
from oracle import rfsi_oracle_np as onp

pts = st.lists(st.tuples(st.integers(0, 6), st.integers(0, 6)), min_size=1, max_size=40)


@settings(max_examples=80, deadline=None)
@given(obs=pts, loc=pts, n=st.integers(1, 8), rm=st.booleans(), seed=st.integers(0, 10))
def test_three_near_obs_restatements_agree(obs, loc, n, rm, seed):
    obs = np.array(obs, dtype=np.float64) * 0.5
    loc = np.array(loc, dtype=np.float64) * 0.5
    z = np.random.default_rng(seed).normal(size=len(obs))
    d1, o1, i1, rc1 = oracle.near_obs(loc, obs, z, n, rm_dupl=rm)
    d2, o2, i2, rc2 = oracle.near_obs(loc, obs, z, n, rm_dupl=rm, kdtree=True, nthreads=2)
    d3, o3, i3 = onp.near_obs(loc, obs, z, n, rm_dupl=rm)
    assert rc1 == rc2
    for a, b in ((d1, d2), (i1, i2), (d1, d3), (i1, i3)):
        assert np.array_equal(a, b, equal_nan=True)
    # invariants: distances ascending; every distance is the one to the reported station
    ok = i1 != onp.NA_INTEGER
    dd = np.where(ok, d1, np.inf)
    assert np.all(dd[:, 1:] >= dd[:, :-1])
    for p in range(len(loc)):
        for j in range(n):
            if ok[p, j]:
                s = i1[p, j] - 1
                assert d1[p, j] == np.sqrt((loc[p, 0] - obs[s, 0]) ** 2 + (loc[p, 1] - obs[s, 1]) ** 2)
                assert o1[p, j] == z[s]
    # with rm_dupl the first neighbour is never a zero-distance one unless a second coincident station exists
    if rm:
        for p in range(len(loc)):
            if ok[p, 0] and d1[p, 0] == 0.0:
                assert (np.abs(obs - loc[p]).sum(axis=1) == 0).sum() >= 2


@settings(max_examples=60, deadline=None)
@given(vals=st.lists(st.one_of(st.floats(-5, 5, allow_nan=False), st.just(float("nan")), st.just(1.25)), min_size=1,
                     max_size=60),
       p=st.floats(0, 1))
def test_quantile7_matches_numpy_linear(vals, p):
    v = np.array(vals)
    got = onp.quantile7(v, [p])[0]
    clean = np.sort(v[~np.isnan(v)])
    if len(clean) == 0:
        assert np.isnan(got)
        return
    assert got == oracle.quantile7_sorted(clean, p)
    assert abs(got - np.quantile(clean, p)) <= 1e-12 * max(1.0, abs(got))
    assert clean[0] <= got <= clean[-1]
