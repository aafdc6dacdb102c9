"""Shared helpers for the GPU parity tests."""
import numpy as np

RTOL = 1e-12   # north_star: predictions within 1e-12 relative in fp64


def same_mean(got, exp):
    """Forest means: within the north_star tolerance, and -- because the engine adds leaf values
    in sequential tree order like ranger and the oracle -- bit-equal."""
    got, exp = np.asarray(got), np.asarray(exp)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
    ok = ~np.isnan(exp)
    np.testing.assert_allclose(got[ok], exp[ok], rtol=RTOL, atol=0)
    np.testing.assert_array_equal(got[ok], exp[ok])


def reference_idw_cv(near_obs_days):
    """The reference's stored IDW cross-validation predictions (temp_croatia/1_models.R:405-536,
    temp_data/IDW_10f_pred.rda; gstat, `nmax = n, idp = p`, UTM33 coordinates) recomputed from a
    near.obs implementation: an IDW prediction is sum(z_i d_i^-p) / sum(d_i^-p) over exactly the
    n nearest stations with an observation that day -- near.obs's output.

    near_obs_days(obs_xy, obs_z[ndays, nobs], n, locations) -> (dist, obs), each (ndays, nloc, n).
    Returns (recomputed, stored), both in the stored order: fold by fold, day by day, station by
    station.  Fixture + the recovery of the per-fold (n, p): tests/golden/make_golden.py."""
    from pathlib import Path
    gold = Path(__file__).parent / "golden"
    g, a = np.load(gold / "croatia_llocv.npz"), np.load(gold / "croatia_idw_cv.npz")
    xy, temp, fold = g["xy"], g["TEMP"], g["fold"]
    out = []
    for k in range(1, 11):
        val, dev = np.flatnonzero(fold == k), np.flatnonzero(fold != k)
        n, p = int(a["n"][k - 1]), float(a["p"][k - 1])
        dist, obs = near_obs_days(xy[dev], temp[:, dev], n, xy[val])
        w = dist ** -p
        pred = (w * obs).sum(axis=2) / w.sum(axis=2)              # (ndays, nval)
        out.append(pred[~np.isnan(temp[:, val])])                 # the val stations observed that day, day-major
    return np.concatenate(out), a["pred"]
