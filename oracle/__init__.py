"""ctypes access to the CPU checker (oracle/rfsi_oracle.c, oracle/cpu_baseline.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product (rfsi_b200/) never
imports this package.  near.obs is pinned against the reference's stored IDW cross-validation
vectors, the forest half is PARITY UNPINNED -- see the header of rfsi_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_P = C.c_void_p
_libs = {}


def build(force: bool = False) -> None:
    if force or not ((_HERE / "liboracle.so").exists() and (_HERE / "libcpubaseline.so").exists()):
        subprocess.run(["make", "-C", str(_HERE)] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)


def _lib(name: str) -> C.CDLL:
    if name not in _libs:
        build()
        _libs[name] = C.CDLL(str(_HERE / name))
    return _libs[name]


def _p(a):
    return None if a is None else a.ctypes.data_as(_P)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def hardware_threads() -> int:
    return int(_lib("libcpubaseline.so").cpu_hardware_threads())


def near_obs(loc_xy, obs_xy, obs_z, n_obs, rm_dupl=True, kdtree=False, nthreads=1):
    """-> (dist[npix,n], obs[npix,n], idx[npix,n] 1-based, rc)."""
    loc_xy = np.asarray(loc_xy, dtype=np.float64)
    obs_xy = np.asarray(obs_xy, dtype=np.float64).reshape(-1, 2)
    lx, ly = _f64(loc_xy[:, 0]), _f64(loc_xy[:, 1])
    ox, oy, oz = _f64(obs_xy[:, 0]), _f64(obs_xy[:, 1]), _f64(obs_z)
    npix, nobs = lx.shape[0], ox.shape[0]
    dist = np.empty((n_obs, npix)); obs = np.empty((n_obs, npix)); idx = np.empty((n_obs, npix), dtype=np.int32)
    if kdtree:
        fn = _lib("libcpubaseline.so").cpu_near_obs_kdtree
        fn.restype = C.c_int
        fn.argtypes = [_P, _P, C.c_int64, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, C.c_int]
        rc = fn(_p(lx), _p(ly), npix, _p(ox), _p(oy), _p(oz), nobs, n_obs, int(rm_dupl), _p(dist), _p(obs), _p(idx),
                int(nthreads))
    else:
        fn = _lib("liboracle.so").oracle_near_obs
        fn.restype = C.c_int
        fn.argtypes = [_P, _P, C.c_int64, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]
        rc = fn(_p(lx), _p(ly), npix, _p(ox), _p(oy), _p(oz), nobs, n_obs, int(rm_dupl), _p(dist), _p(obs), _p(idx))
    if rc < 0:
        raise RuntimeError(f"oracle near_obs failed: {rc}")
    return dist.T, obs.T, idx.T, rc


def _forest_args(fo):
    return (int(fo["ntrees"]), np.ascontiguousarray(fo["node_offsets"], dtype=np.int64),
            np.ascontiguousarray(fo["left"], dtype=np.int32), np.ascontiguousarray(fo["right"], dtype=np.int32),
            np.ascontiguousarray(fo["split_var"], dtype=np.int32), _f64(fo["split_val"]))


def _xt(X):
    """(npix, F) -> column-major buffer, (F, npix) C-order."""
    X = np.asarray(X, dtype=np.float64)
    return np.ascontiguousarray(X.T), X.shape[0]


def forest_predict(fo, X, threads: int = 0, count_visits: bool = False):
    T, off, L, R, V, S = _forest_args(fo)
    Xt, npix = _xt(X)
    out = np.empty(npix)
    if threads > 0:
        fn = _lib("libcpubaseline.so").cpu_forest_predict
        fn.restype = C.c_int
        fn.argtypes = [C.c_int32, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P, C.c_int]
        rc = fn(T, _p(off), _p(L), _p(R), _p(V), _p(S), _p(Xt), npix, npix, _p(out), int(threads))
        visits = None
    else:
        fn = _lib("liboracle.so").oracle_forest_predict
        fn.restype = C.c_int
        fn.argtypes = [C.c_int32, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P, _P]
        v = C.c_int64(0)
        rc = fn(T, _p(off), _p(L), _p(R), _p(V), _p(S), _p(Xt), npix, npix, _p(out),
                C.cast(C.byref(v), _P) if count_visits else None)
        visits = v.value
    if rc < 0:
        raise RuntimeError(f"oracle forest_predict failed: {rc}")
    return (out, visits) if count_visits else out


def forest_terminal_nodes(fo, X):
    T, off, L, R, V, S = _forest_args(fo)
    Xt, npix = _xt(X)
    out = np.empty((T, npix), dtype=np.int32)
    fn = _lib("liboracle.so").oracle_forest_terminal_nodes
    fn.restype = C.c_int
    fn.argtypes = [C.c_int32, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P]
    rc = fn(T, _p(off), _p(L), _p(R), _p(V), _p(S), _p(Xt), npix, npix, _p(out))
    if rc < 0:
        raise RuntimeError(f"oracle terminal_nodes failed: {rc}")
    return out.T


def forest_quantiles(fo, X, probs, threads: int = 0):
    T, off, L, R, V, S = _forest_args(fo)
    Xt, npix = _xt(X)
    probs = _f64(probs)
    roff = np.ascontiguousarray(fo["rnv_offsets"], dtype=np.int64)
    rnv = _f64(fo["random_node_values"])
    out = np.empty((len(probs), npix))
    if threads > 0:
        fn = _lib("libcpubaseline.so").cpu_forest_quantiles
        fn.restype = C.c_int
        fn.argtypes = [C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P, C.c_int32, _P, C.c_int]
        rc = fn(T, _p(off), _p(L), _p(R), _p(V), _p(S), _p(roff), _p(rnv), _p(Xt), npix, npix, _p(probs), len(probs),
                _p(out), int(threads))
    else:
        fn = _lib("liboracle.so").oracle_forest_quantiles
        fn.restype = C.c_int
        fn.argtypes = [C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P, C.c_int32, _P]
        rc = fn(T, _p(off), _p(L), _p(R), _p(V), _p(S), _p(roff), _p(rnv), _p(Xt), npix, npix, _p(probs), len(probs),
                _p(out))
    if rc < 0:
        raise RuntimeError(f"oracle forest_quantiles failed: {rc}")
    return out.T


def quantile7_sorted(xs, p) -> float:
    xs = _f64(xs)
    fn = _lib("liboracle.so").oracle_quantile7_sorted
    fn.restype = C.c_double
    fn.argtypes = [_P, C.c_int32, C.c_double]
    return float(fn(_p(xs), len(xs), float(p)))


def centi_int16(x):
    x = _f64(x).ravel()
    out = np.empty(x.shape[0], dtype=np.int16)
    fn = _lib("liboracle.so").oracle_centi_int16
    fn.restype = C.c_int
    fn.argtypes = [_P, C.c_int64, _P]
    fn(_p(x), x.shape[0], _p(out))
    return out


def predict_fused(fo, feature_names, *, loc_xy, obs_xy, obs_z, n_obs, covariates=None, quantiles=None,
                  rm_dupl=True, kdtree=False, threads=0):
    """What the mapping loops do per day (2_predictions.R:122-125,160-168,174-178):
    keep that day's stations (!is.na), near.obs, cbind covariates, predict.
    -> {"mean": (ndays, npix), "quantiles": (Q, ndays, npix)}; pixel-days with fewer
    valid stations than needed are NA."""
    from .rfsi_oracle_np import NA_REAL
    loc_xy = np.asarray(loc_xy, dtype=np.float64)
    obs_xy = np.asarray(obs_xy, dtype=np.float64)
    obs_z = np.atleast_2d(np.asarray(obs_z, dtype=np.float64))
    covariates = covariates or {}
    ndays, npix = obs_z.shape[0], loc_xy.shape[0]
    mean = np.full((ndays, npix), NA_REAL)
    q = None if quantiles is None else np.full((len(quantiles), ndays, npix), NA_REAL)
    for d in range(ndays):
        valid = ~np.isnan(obs_z[d])
        dist, obs, _, _ = near_obs(loc_xy, obs_xy[valid], obs_z[d][valid], n_obs, rm_dupl, kdtree=kdtree,
                                   nthreads=max(threads, 1))
        cols = []
        for name in feature_names:
            if name.startswith("dist") and name[4:].isdigit():
                cols.append(dist[:, int(name[4:]) - 1])
            elif name.startswith("obs") and name[3:].isdigit():
                cols.append(obs[:, int(name[3:]) - 1])
            else:
                c = np.asarray(covariates[name], dtype=np.float64)
                cols.append(c if c.ndim == 1 else c[d])
        X = np.stack(cols, axis=1)
        good = ~np.isnan(X).any(axis=1)
        if good.any():
            mean[d, good] = forest_predict(fo, X[good], threads=threads)
            if q is not None:
                q[:, d, good] = forest_quantiles(fo, X[good], quantiles, threads=threads).T
    return {"mean": mean, "quantiles": q}
