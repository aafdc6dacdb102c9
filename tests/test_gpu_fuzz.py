"""Randomised small cases on a real B200 against the CPU oracle: odd sizes, lattices full of
ties, coincident stations, short station lists, NA observations, random forests with features
sitting exactly on thresholds, random day/mask patterns.  Everything must be bit-equal."""
import numpy as np
import pytest

import oracle
import rfsi_b200 as rb
from rfsi_b200.forest import FlatForest
from tests.util import same_mean

pytestmark = pytest.mark.gpu


def _rand_points(rng, n, lattice):
    if lattice:
        return rng.integers(0, 12, size=(n, 2)).astype(np.float64) * 0.5          # many ties / coincident points
    return rng.uniform(-50, 50, size=(n, 2))


def _knn_gpu(loc, obs, z, n, rm_dupl):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df, idx = rb.near_obs(loc, np.column_stack([obs, z]), zcol=2, n_obs=n, rm_dupl=rm_dupl, return_idx=True)
    a = df.to_numpy()
    return a[:, :n], a[:, n:], idx


@pytest.mark.parametrize("cells_min", ["128", "1"])
def test_fuzz_near_obs(cells_min, monkeypatch):
    monkeypatch.setenv("RFSI_B200_CELLS_MIN", cells_min)
    rng = np.random.default_rng(2024)
    for it in range(120):
        S = int(rng.choice([1, 2, 3, 5, 17, 64, 130, 300]))
        npix = int(rng.integers(1, 400))
        n = int(rng.integers(1, 14))
        lattice = bool(rng.integers(0, 2))
        obs, loc = _rand_points(rng, S, lattice), _rand_points(rng, npix, lattice)
        z = rng.normal(size=S)
        if rng.uniform() < 0.3:
            z[rng.integers(0, S)] = np.nan                       # NA observation values propagate
        rm = bool(rng.integers(0, 2))
        d, o, i = _knn_gpu(loc, obs, z, n, rm)
        ed, eo, ei, _ = oracle.near_obs(loc, obs, z, n, rm_dupl=rm)
        assert np.array_equal(i, ei), (it, S, npix, n, lattice, rm)
        assert np.array_equal(d.view(np.uint64), ed.view(np.uint64)), (it, S, npix, n)
        assert np.array_equal(np.isnan(o), np.isnan(eo)) and np.array_equal(o[~np.isnan(eo)], eo[~np.isnan(eo)])


def _random_forest(rng, T, F, max_nodes, quantreg=True):
    """Random ranger-layout trees: random shapes (incl. single-leaf trees and long chains), split
    values drawn from a small grid so features can sit exactly on thresholds."""
    left, right, var, val, off, rnv, roff = [], [], [], [], [0], [], []
    for _ in range(T):
        n_int = int(rng.integers(0, max_nodes))
        l, r, v, s = [0], [0], [0], [0.0]
        frontier = [0]
        for _k in range(n_int):
            node = frontier.pop(int(rng.integers(0, len(frontier))) if rng.uniform() < 0.7 else -1)
            l[node], r[node] = len(l), len(l) + 1
            v[node] = int(rng.integers(0, F)); s[node] = float(rng.integers(-4, 5)) * 0.5
            for _c in range(2):
                l.append(0); r.append(0); v.append(0); s.append(0.0)
            frontier += [len(l) - 2, len(l) - 1]
        for node in range(len(l)):
            if l[node] == 0:
                s[node] = float(rng.normal())
        left += l; right += r; var += v; val += s
        roff.append(len(rnv))
        samp = rng.normal(size=len(l)); samp[rng.uniform(size=len(l)) < 0.15] = np.nan
        samp[rng.uniform(size=len(l)) < 0.2] = 0.25                # duplicates
        rnv += list(samp)
        off.append(len(left))
    return FlatForest(T, np.array(off, dtype=np.int64), np.array(left, dtype=np.int32), np.array(right, dtype=np.int32),
                      np.array(var, dtype=np.int32), np.array(val), [f"f{i}" for i in range(F)],
                      np.array(roff, dtype=np.int64) if quantreg else None, np.array(rnv) if quantreg else None)


@pytest.mark.parametrize("fmt", ["3", "2", "1", "0"])
def test_fuzz_forest(fmt, monkeypatch):
    monkeypatch.setenv("RFSI_B200_COMPACT", fmt)
    rng = np.random.default_rng(7 + int(fmt))
    for it in range(40):
        T, F = int(rng.integers(1, 40)), int(rng.integers(1, 9))
        flat = _random_forest(rng, T, F, int(rng.choice([1, 3, 20, 200])))
        model = rb.RangerForest(flat)
        npix = int(rng.integers(1, 700))
        X = rng.integers(-5, 6, size=(npix, F)).astype(np.float64) * 0.5      # on the threshold grid
        if rng.uniform() < 0.5:
            X += rng.normal(scale=0.3, size=X.shape)
        fo = flat.as_dict()
        same_mean(rb.predict(model, X).predictions, oracle.forest_predict(fo, X))
        assert np.array_equal(rb.predict(model, X, type="terminalNodes").predictions, oracle.forest_terminal_nodes(fo, X)), it
        probs = sorted(rng.uniform(size=int(rng.integers(1, 5))).tolist()) + [0.0, 1.0]
        gq = rb.predict(model, X, type="quantiles", quantiles=probs).predictions
        eq = oracle.forest_quantiles(fo, X, probs)
        assert np.array_equal(np.isnan(gq), np.isnan(eq)) and np.array_equal(gq[~np.isnan(eq)], eq[~np.isnan(eq)]), it
        model.close()


def test_fuzz_fused():
    rng = np.random.default_rng(99)
    for it in range(30):
        S, D, n = int(rng.choice([8, 40, 150])), int(rng.integers(1, 5)), int(rng.integers(1, 7))
        lattice = bool(rng.integers(0, 2))
        obs = _rand_points(rng, S, lattice)
        z = rng.normal(size=(D, S))
        z[rng.uniform(size=z.shape) < rng.choice([0.0, 0.1, 0.6])] = np.nan
        names = ["c0"] + [f"dist{j + 1}" for j in range(n)] + [f"obs{j + 1}" for j in range(n)]
        order = rng.permutation(len(names))
        names = [names[k] for k in order]                                   # model columns in any order
        flat = _random_forest(rng, int(rng.integers(1, 30)), len(names), 60)
        flat.feature_names = names
        model = rb.RangerForest(flat)
        ncol, nrow = int(rng.integers(1, 40)), int(rng.integers(1, 30))
        npix = ncol * nrow
        grid = (0.25, 0.25, 0.5, 0.5, ncol) if lattice else (-50.0, -50.0, 100.0 / ncol, 100.0 / nrow, ncol)
        p = np.arange(npix)
        loc = np.stack([grid[0] + (p % ncol) * grid[2], grid[1] + (p // ncol) * grid[3]], axis=1)
        cov = {"c0": rng.integers(-3, 4, size=(D, npix)).astype(np.float64) * 0.5}
        mask = (rng.uniform(size=npix) < 0.8).astype(np.uint8) if rng.uniform() < 0.5 else None
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            use_grid = bool(rng.integers(0, 2))
            kw = dict(grid=grid, npix=npix) if use_grid else dict(locations=loc)
            got = rb.predict_fused(model, obs_xy=obs, obs_z=z, n_obs=n, covariates=cov, quantiles=[0.25, 0.5, 0.75],
                                   pix_mask=mask, **kw)
        ref = oracle.predict_fused(flat.as_dict(), names, loc_xy=loc, obs_xy=obs, obs_z=z, n_obs=n, covariates=cov,
                                   quantiles=[0.25, 0.5, 0.75])
        if mask is not None:
            ref["mean"][:, mask == 0] = np.nan; ref["quantiles"][:, :, mask == 0] = np.nan
        same_mean(got["mean"], ref["mean"])
        ok = ~np.isnan(ref["quantiles"])
        assert np.array_equal(np.isnan(got["quantiles"]), ~ok), it
        assert np.array_equal(got["quantiles"][ok], ref["quantiles"][ok]), it
        model.close()
