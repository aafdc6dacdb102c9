"""NumPy restatement of the RFSI prediction hot path (readable twin of rfsi_oracle.c).

TEST INFRASTRUCTURE ONLY -- only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this; the product (rfsi_b200/) never does.

PINNING: stage 1 (near.obs) IS pinned against an output of the reference itself -- the 55 746
stored predictions of its nested 10-fold IDW cross-validation (temp_croatia/1_models.R:405-536,
temp_data/IDW_10f_pred.rda, computed in R by gstat), each a function of exactly near.obs's output
(the n nearest observed stations, their distances and values), are reproduced to 2e-14
(tests/test_oracle.py, tests/util.py::reference_idw_cv).  Stage 2 (ranger predict / quantiles) is
PARITY UNPINNED: the reference holds no saved forest or golden vectors for it and its
arithmetic lives in un-vendored R packages that cannot run here.
Semantics follow SURVEY.md Appendix A, anchored on the reference call sites:
near.obs -> simulation/simulation.R:191-196; predict -> simulation.R:214;
quantiles -> temp_croatia/2_predictions.R:174-178.
"""
from __future__ import annotations

import numpy as np

NA_REAL = np.frombuffer(np.uint64(0x7FF00000000007A2).tobytes(), dtype=np.float64)[0]
NA_INTEGER = np.int32(-2147483648)


def near_obs(loc_xy, obs_xy, obs_z, n_obs, rm_dupl=True):
    """near.obs: (dist[npix,n], obs[npix,n], idx[npix,n] 1-based).

    Total order (d2, index): a stable argsort of d2 keeps the lower index first
    among equal squared distances, which is what an ascending-index scan with a
    strict '<' insertion test produces (Appendix A.1 item 7).
    """
    loc_xy = np.asarray(loc_xy, dtype=np.float64)
    obs_xy = np.asarray(obs_xy, dtype=np.float64)
    obs_z = np.asarray(obs_z, dtype=np.float64)
    npix, nobs = loc_xy.shape[0], obs_xy.shape[0]
    k = n_obs + 1
    dist = np.full((npix, n_obs), NA_REAL)
    obs = np.full((npix, n_obs), NA_REAL)
    idx = np.full((npix, n_obs), NA_INTEGER, dtype=np.int32)
    for p in range(npix):
        dx = loc_xy[p, 0] - obs_xy[:, 0]
        dy = loc_xy[p, 1] - obs_xy[:, 1]
        d2 = dx * dx + dy * dy          # two roundings for the squares, one for the sum
        order = np.argsort(d2, kind="stable")[:k]
        first = 1 if (rm_dupl and nobs > 0 and d2[order[0]] == 0.0) else 0
        keep = order[first:first + n_obs]
        m = keep.shape[0]
        dist[p, :m] = np.sqrt(d2[keep])
        obs[p, :m] = obs_z[keep]
        idx[p, :m] = keep + 1
    return dist, obs, idx


def _descend(left, right, svar, sval, x):
    node = 0
    visits = 0
    while left[node] != 0 or right[node] != 0:
        node = left[node] if x[svar[node]] <= sval[node] else right[node]
        visits += 1
    return node, visits


def forest_terminal_nodes(forest, X):
    """forest: dict(ntrees, node_offsets, left, right, split_var, split_val). X[npix,F]."""
    X = np.asarray(X, dtype=np.float64)
    T = int(forest["ntrees"])
    out = np.empty((X.shape[0], T), dtype=np.int32)
    off = forest["node_offsets"]
    for t in range(T):
        a, b = int(off[t]), int(off[t + 1])
        L, R = forest["left"][a:b], forest["right"][a:b]
        V, S = forest["split_var"][a:b], forest["split_val"][a:b]
        for r in range(X.shape[0]):
            out[r, t] = _descend(L, R, V, S, X[r])[0]
    return out


def forest_predict(forest, X):
    """(sum over trees in tree order of leaf values) / ntrees (Appendix A.2)."""
    X = np.asarray(X, dtype=np.float64)
    T = int(forest["ntrees"])
    off = forest["node_offsets"]
    tn = forest_terminal_nodes(forest, X)
    out = np.zeros(X.shape[0])
    for t in range(T):
        out = out + forest["split_val"][int(off[t]) + tn[:, t]]
    return out / float(T)


def quantile7(values, probs):
    """R quantile(x, probs, na.rm=TRUE, type=7) (Appendix A.3 item 3)."""
    v = np.asarray(values, dtype=np.float64)
    v = np.sort(v[~np.isnan(v)])
    m = v.shape[0]
    out = np.empty(len(probs))
    for i, p in enumerate(probs):
        if m == 0:
            out[i] = NA_REAL
            continue
        index = 1.0 + (m - 1) * p
        lo, hi = np.floor(index), np.ceil(index)
        xlo, xhi = v[int(lo) - 1], v[int(hi) - 1]
        if index > lo and xhi != xlo:
            h = index - lo
            out[i] = (1.0 - h) * xlo + h * xhi
        else:
            out[i] = xlo
    return out


def forest_quantiles(forest, X, probs):
    """predict(type='quantiles'): random_node_values[leaf, tree] -> type-7 quantiles."""
    tn = forest_terminal_nodes(forest, X)
    roff = forest["rnv_offsets"]
    rnv = forest["random_node_values"]
    out = np.empty((tn.shape[0], len(probs)))
    for r in range(tn.shape[0]):
        vals = np.array([rnv[int(roff[t]) + tn[r, t]] for t in range(tn.shape[1])])
        out[r] = quantile7(vals, probs)
    return out


def raster_extract(data, x_ul, y_ul, dx, dy, nodata, xy, layer=0, scale=1.0, offset=0.0, divide=False):
    """raster::extract(r, points), method "simple", plus the unit change the scripts apply to it
    (prcp_catalonia/5_prcp_prediction.R:253-276: `extract(tmin, gg) / 10`).

    The value of the cell each point falls in: col = floor((x - xmin)/xres), row counted from
    the top, the right and bottom outer edges belong to the last column / row
    (raster::cellFromXY); NA outside the raster and where the cell holds the NAvalue.
    UNPINNED against the raster package (R cannot run here); the arithmetic is two IEEE
    operations, `raw / scale` (or `raw * scale`) then `+ offset`."""
    d = data if data.ndim == 3 else data[None]
    nrow, ncol = d.shape[1:]
    xy = np.asarray(xy, dtype=np.float64)
    x, y = xy[:, 0], xy[:, 1]
    inside = (x >= x_ul) & (x <= x_ul + ncol * dx) & (y <= y_ul) & (y >= y_ul - nrow * dy)
    with np.errstate(invalid="ignore"):
        c = np.clip(np.nan_to_num(np.minimum(np.floor((x - x_ul) / dx), ncol - 1)), 0, ncol - 1).astype(np.int64)
        r = np.clip(np.nan_to_num(np.minimum(np.floor((y_ul - y) / dy), nrow - 1)), 0, nrow - 1).astype(np.int64)
    raw = d[layer][r, c].astype(np.float64)
    ok = inside & ~(raw == nodata) & ~np.isnan(raw)
    out = np.full(len(x), NA_REAL)
    out[ok] = (raw[ok] / scale if divide else raw[ok] * scale) + offset
    return out


def cov_fix(target, other, delta):
    """`df$tmin <- ifelse(df$tmin > df$tmax, df$tmax-1, df$tmin)` (5_prcp_prediction.R:258):
    NA where either side is NA (R's ifelse on an NA test)."""
    target = np.asarray(target, dtype=np.float64)
    other = np.asarray(other, dtype=np.float64)
    na = np.isnan(target) | np.isnan(other)
    with np.errstate(invalid="ignore"):
        out = np.where(target > other, other - delta, target)
    out[na] = NA_REAL
    return out
