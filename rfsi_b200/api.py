"""Host-side mirror of the reference's R call shapes for the hot path.

R is not installed in this image, so the host side above the C ABI is Python with the
reference's names, argument meaning and error behaviour (R's dotted argument names
become underscores: n.obs -> n_obs).  The R shim a maintainer would add is under
rfsi_b200/R and documented in INTEGRATION.md.

  near_obs(locations, observations, zcol, n_obs)   simulation/simulation.R:191-196
  predict(model, data)                              simulation/simulation.R:214
  predict(model, data, type="quantiles", quantiles) temp_croatia/2_predictions.R:174-176
  rfsi(...), pred_rfsi(...)                         README.md:93-143

Every compute call goes through librfsi_b200.so (CUDA, sm_100a); nothing here computes
neighbours, tree descents or quantiles on the CPU.
"""
from __future__ import annotations

import ctypes as C
import warnings
from typing import Optional, Sequence

import numpy as np

from . import _lib
from dataclasses import dataclass

from ._lib import FEAT_DIST, FEAT_OBS, RASTER_DIVIDE, RASTER_DTYPES, CovFix, FusedArgs, Raster as _CRaster, RfsiError
from .forest import FlatForest, RangerForest, flatten_ranger, from_sklearn

try:  # pandas plays the role of R's data.frame
    import pandas as pd
except Exception:  # pragma: no cover
    pd = None


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _coords(obj, x_y) -> np.ndarray:
    """coordinates(Spatial*) or data.frame[, locations.x.y]  (Appendix A.1 item 1)."""
    if pd is not None and isinstance(obj, pd.DataFrame):
        cols = [obj.columns[c] if isinstance(c, (int, np.integer)) else c for c in x_y]
        return obj[cols].to_numpy(dtype=np.float64)
    a = np.asarray(obj, dtype=np.float64)
    if a.ndim != 2 or a.shape[1] < 2:
        raise ValueError("locations/observations must be a data frame or an (n x >=2) array")
    return a[:, [int(x_y[0]), int(x_y[1])]]


def _column(obj, col) -> np.ndarray:
    if pd is not None and isinstance(obj, pd.DataFrame):
        name = obj.columns[col] if isinstance(col, (int, np.integer)) else col
        return obj[name].to_numpy(dtype=np.float64)
    return np.asarray(obj, dtype=np.float64)[:, int(col)]


def near_obs(locations, observations, zcol=2, n_obs: int = 10, locations_x_y=(0, 1), observations_x_y=(0, 1),
             rm_dupl: bool = True, avg: bool = False, direct: bool = False, idw: bool = False,
             return_idx: bool = False, device: int = 0, precision: str = "fp64"):
    """meteo::near.obs: for every location the n_obs nearest observations.

    Returns a data frame with columns dist1..distN then obs1..obsN (the order
    sim_500.R:146 relies on; names relied on at temp_croatia/1_models.R:978,1053).
    Column indices are 0-based here (R's c(1,2) default becomes (0, 1), zcol=3 -> 2).
    With return_idx also returns the 1-based neighbour indices (npix x n_obs).
    precision="fp32": fp32 mode (rfsi_near_obs_f32) -- coordinates and values are rounded to float first, float
    buffers cross PCIe, fp32 distances prefilter the search and the survivors are re-ranked in fp64: neighbours are
    those of the fp64 call on the rounded coordinates, dist/obs come back as float32 columns.
    """
    if avg or direct or idw:
        # the reference passes only locations, observations, zcol, n.obs (all 20 call sites)
        raise NotImplementedError("near_obs: avg/direct/idw are not supported")
    n_obs = int(n_obs)
    if n_obs < 1:
        raise ValueError("n.obs must be >= 1")
    loc = _coords(locations, locations_x_y)
    obs = _coords(observations, observations_x_y)
    z = _column(observations, zcol)
    npix, nobs = loc.shape[0], obs.shape[0]
    if precision not in ("fp64", "fp32"):
        raise ValueError("precision must be 'fp64' or 'fp32'")
    names = [f"dist{j + 1}" for j in range(n_obs)] + [f"obs{j + 1}" for j in range(n_obs)]
    idx = np.empty((n_obs, npix), dtype=np.int32) if return_idx else None
    if precision == "fp32":
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        lx, ly, ox, oy, oz = f32(loc[:, 0]), f32(loc[:, 1]), f32(obs[:, 0]), f32(obs[:, 1]), f32(z)
        both = np.empty((2, n_obs, npix), dtype=np.float32)
        rc = _lib.load().rfsi_near_obs_f32(_ptr(lx), _ptr(ly), npix, _ptr(ox), _ptr(oy), _ptr(oz), nobs, n_obs,
                                          1 if rm_dupl else 0, _ptr(both[0]), _ptr(both[1]), _ptr(idx), device)
        if _lib.check(rc) == _lib.RFSI_W_TOO_FEW_OBS:
            warnings.warn(_lib.last_error())
        cols = [both[0, j] for j in range(n_obs)] + [both[1, j] for j in range(n_obs)]
        out = pd.DataFrame(dict(zip(names, cols)), copy=False) if pd is not None else dict(zip(names, cols))
        return (out, idx.T) if return_idx else out
    lx, ly = _f64(loc[:, 0]), _f64(loc[:, 1])
    ox, oy, oz = _f64(obs[:, 0]), _f64(obs[:, 1]), _f64(z)
    # one array per column, filled in place by the engine (rfsi_near_obs_cols_f64): the frame below wraps them
    cols = [np.empty(npix, dtype=np.float64) for _ in range(2 * n_obs)]
    cptr = (C.c_void_p * (2 * n_obs))(*[c.ctypes.data for c in cols])
    iptr = (C.c_void_p * n_obs)(*[idx[j].ctypes.data for j in range(n_obs)]) if return_idx else None
    rc = _lib.load().rfsi_near_obs_cols_f64(_ptr(lx), _ptr(ly), npix, _ptr(ox), _ptr(oy), _ptr(oz), nobs, n_obs,
                                           1 if rm_dupl else 0, cptr, C.byref(cptr, n_obs * C.sizeof(C.c_void_p)), iptr, device)
    if _lib.check(rc) == _lib.RFSI_W_TOO_FEW_OBS:
        warnings.warn(_lib.last_error())
    out = pd.DataFrame(dict(zip(names, cols)), copy=False) if pd is not None else dict(zip(names, cols))
    return (out, idx.T) if return_idx else out


def near_obs_days(obs_xy, obs_z, n_obs: int, locations=None, rm_dupl: bool = True, return_idx: bool = False,
                  device: int = 0):
    """near.obs for every day of a station table in one call (rfsi_near_obs_days_f64).

    obs_z: (ndays, nobs), NaN = no observation that day.  locations=None: the stations
    themselves (the training-table shape of temp_croatia/1_models.R:1034-1047: each row's
    own observation is excluded; rows of silent stations are NA).
    Returns dist, obs of shape (ndays, npix, n_obs) (+ idx, 1-based original station index).
    """
    obs_xy = np.asarray(obs_xy, dtype=np.float64)
    obs_z = _f64(np.atleast_2d(obs_z))
    ndays, nobs = obs_z.shape
    if obs_xy.shape[0] != nobs:
        raise ValueError("obs_z must be (ndays, nobs) with nobs == len(obs_xy)")
    ox, oy = _f64(obs_xy[:, 0]), _f64(obs_xy[:, 1])
    lx = ly = None
    npix = nobs
    if locations is not None:
        loc = np.asarray(locations, dtype=np.float64)
        lx, ly = _f64(loc[:, 0]), _f64(loc[:, 1])
        npix = loc.shape[0]
    dist = np.empty((ndays, n_obs, npix)); val = np.empty((ndays, n_obs, npix))
    idx = np.empty((ndays, n_obs, npix), dtype=np.int32) if return_idx else None
    rc = _lib.load().rfsi_near_obs_days_f64(_ptr(lx), _ptr(ly), npix, _ptr(ox), _ptr(oy), _ptr(obs_z), nobs, ndays,
                                           int(n_obs), 1 if rm_dupl else 0, _ptr(dist), _ptr(val), _ptr(idx), device)
    _lib.check(rc)          # too-few rows are simply NA here (complete.cases drops them, 1_models.R:1003)
    out = (dist.transpose(0, 2, 1), val.transpose(0, 2, 1))
    return out + (idx.transpose(0, 2, 1),) if return_idx else out


class Prediction:
    """What predict.ranger returns: a list with $predictions (vector or npix x Q matrix)."""

    def __init__(self, predictions: np.ndarray, num_trees: int, quantiles=None):
        self.predictions = predictions
        self.num_trees = num_trees
        self.quantiles = quantiles
        self.treetype = "Regression"
        self.num_samples = predictions.shape[0]


def _model_matrix(model: RangerForest, data) -> np.ndarray:
    """Select / re-order columns by independent.variable.names; extra columns ignored
    (simulation.R:214 passes sim_df incl. sim1, fold, x, y).  Returns F x npix C-order =
    npix x F column-major."""
    names = model.feature_names
    if pd is not None and isinstance(data, pd.DataFrame):
        missing = [n for n in names if n not in data.columns]
        if missing:
            raise ValueError(f"predict: columns missing from data: {missing}")
        return np.ascontiguousarray(np.stack([data[n].to_numpy(dtype=np.float64) for n in names], axis=0))
    if isinstance(data, dict):
        missing = [n for n in names if n not in data]
        if missing:
            raise ValueError(f"predict: columns missing from data: {missing}")
        return np.ascontiguousarray(np.stack([np.asarray(data[n], dtype=np.float64) for n in names], axis=0))
    a = np.asarray(data, dtype=np.float64)
    if a.ndim != 2 or a.shape[1] != len(names):
        raise ValueError(f"predict: matrix input must be npix x {len(names)} in the model's column order")
    return np.ascontiguousarray(a.T)


def predict(model: RangerForest, data, type: str = "response", quantiles: Sequence[float] = (0.1, 0.5, 0.9),
            device: int = 0) -> Prediction:
    """predict.ranger for a regression forest: type = "response" | "quantiles" | "terminalNodes"."""
    Xt = _model_matrix(model, data)            # (F, npix) C-order == npix x F column-major
    F, npix = Xt.shape
    lib = _lib.load()
    if type == "response":
        out = np.empty(npix, dtype=np.float64)
        _lib.check(lib.rfsi_forest_predict(model.handle, _ptr(Xt), npix, npix, _ptr(out), device))
        return Prediction(out, model.num_trees)
    if type == "quantiles":
        if not model.has_quantile_samples:
            raise ValueError("Error: Set quantreg=TRUE in ranger(...) for quantile prediction.")
        probs = _f64(quantiles)
        out = np.empty((len(probs), npix), dtype=np.float64)
        _lib.check(lib.rfsi_forest_quantiles(model.handle, _ptr(Xt), npix, npix, _ptr(probs), len(probs), _ptr(out),
                                             device))
        return Prediction(np.ascontiguousarray(out.T), model.num_trees, quantiles=list(probs))
    if type == "terminalNodes":
        out = np.empty((model.num_trees, npix), dtype=np.int32)
        _lib.check(lib.rfsi_forest_terminal_nodes(model.handle, _ptr(Xt), npix, npix, _ptr(out), device))
        return Prediction(np.ascontiguousarray(out.T), model.num_trees)
    raise ValueError(f"predict: unknown type {type!r}")


@dataclass
class Raster:
    """A north-up covariate raster: data (nrow, ncol) or (nlayers, nrow, ncol) of float64 /
    float32 / int32 / int16 / uint16 / uint8; (x_ul, y_ul) the outer corner of the top-left
    cell; value = raw*scale + offset (divide=True: raw/scale + offset, R's `extract(r, p) / 10`);
    raw == nodata -> NA.  Sampled on the device like
    raster::extract(r, points) (prcp_catalonia/5_prcp_prediction.R:253-276)."""
    data: np.ndarray
    x_ul: float
    y_ul: float
    dx: float
    dy: float
    nodata: float = float("nan")
    scale: float = 1.0
    offset: float = 0.0
    divide: bool = False

    def _c(self):
        d = np.ascontiguousarray(self.data)
        if d.ndim == 2:
            d = d[None]
        if d.dtype.name not in RASTER_DTYPES:
            raise ValueError(f"raster dtype {d.dtype} unsupported (use one of {list(RASTER_DTYPES)})")
        r = _CRaster(d.ctypes.data, RASTER_DTYPES[d.dtype.name] | (RASTER_DIVIDE if self.divide else 0), d.shape[0], d.shape[2], d.shape[1],
                     float(self.x_ul), float(self.y_ul), float(self.dx), float(self.dy), float(self.nodata),
                     float(self.scale), float(self.offset))
        return r, d


def raster_extract(raster: Raster, locations=None, grid=None, pix_begin: int = 0, npix: Optional[int] = None,
                   layer: int = 0, device: int = 0) -> np.ndarray:
    """raster::extract(r, points): the value of the cell each location falls in (NA outside)."""
    r, keep = raster._c()
    lx = ly = None
    g = (0.0, 0.0, 0.0, 0.0, 1)
    if locations is not None:
        loc = np.asarray(locations, dtype=np.float64)
        lx, ly = _f64(loc[:, 0]), _f64(loc[:, 1])
        npix = loc.shape[0]
    else:
        if grid is None or npix is None:
            raise ValueError("give locations, or grid=(x0, y0, dx, dy, ncol) and npix")
        g = grid
    out = np.empty(int(npix), dtype=np.float64)
    rc = _lib.load().rfsi_raster_extract(C.byref(r), int(layer), _ptr(lx), _ptr(ly), int(npix), float(g[0]), float(g[1]),
                                         float(g[2]), float(g[3]), int(g[4]), int(pix_begin), _ptr(out), device)
    _lib.check(rc)
    del keep
    return out


def feature_map(feature_names: Sequence[str], cov_names: Sequence[str], n_obs: int) -> np.ndarray:
    """Model column -> RFSI_FEAT_* code: 'dist{j}' / 'obs{j}' (1-based) or a covariate name."""
    codes = []
    cov_names = list(cov_names)
    for n in feature_names:
        if n.startswith("dist") and n[4:].isdigit() and 1 <= int(n[4:]) <= n_obs:
            codes.append(FEAT_DIST + int(n[4:]) - 1)
        elif n.startswith("obs") and n[3:].isdigit() and 1 <= int(n[3:]) <= n_obs:
            codes.append(FEAT_OBS + int(n[3:]) - 1)
        elif n in cov_names:
            codes.append(cov_names.index(n))
        else:
            raise ValueError(f"model column {n!r} is neither dist/obs 1..{n_obs} nor a supplied covariate")
    return np.asarray(codes, dtype=np.int32)


def predict_fused(model: RangerForest, *, obs_xy, obs_z, n_obs: int, locations=None, grid=None,
                  pix_begin: int = 0, npix: Optional[int] = None, covariates: Optional[dict] = None,
                  quantiles: Optional[Sequence[float]] = None, want_mean: bool = True, rm_dupl: bool = True,
                  centi_int16: bool = False, pix_mask=None, cov_fix: Sequence = (), cov_locations=None, device: int = 0,
                  devices: Optional[Sequence[int]] = None) -> dict:
    """near.obs + cbind + predict for every (pixel, day) in one call (rfsi_predict_fused).

    obs_z: (ndays, nobs) observed values, NaN = no observation that day (the scripts'
    per-day `st[!is.na(st$TEMP),]` filter, 2_predictions.R:122-125).
    locations: (npix, 2) coordinates, or grid=(x0, y0, dx, dy, ncol) + npix (+ pix_begin)
    for raster cells in row-major order.  covariates: {name: (npix,) or (ndays, npix) or Raster}.
    cov_fix: [(target, other, delta), ...] covariate repairs applied on the device after sampling,
    `df$target <- ifelse(df$target > df$other, df$other - delta, df$target)`
    (5_prcp_prediction.R:258: ("tmin", "tmax", 1)).  cov_locations: (npix, 2) coordinates at which
    Raster covariates are sampled when they differ from the near.obs coordinates (lon/lat `gg` vs
    projected `gg2`, 5_prcp_prediction.R:42-46).  devices: several GPUs of the box, one band of
    pixels each (rfsi_predict_fused_multi); default the single `device`.
    Returns {"mean": (ndays, npix), "quantiles": (Q, ndays, npix), "mean_i16", "iqr_i16"}.
    """
    lib = _lib.load()
    obs_xy = np.asarray(obs_xy, dtype=np.float64)
    obs_z = _f64(np.atleast_2d(obs_z))
    ndays, nobs = obs_z.shape
    if obs_xy.shape[0] != nobs:
        raise ValueError("obs_z must be (ndays, nobs) with nobs == len(obs_xy)")
    ox, oy = _f64(obs_xy[:, 0]), _f64(obs_xy[:, 1])
    covariates = covariates or {}
    a = FusedArgs()
    a.struct_size = C.sizeof(FusedArgs)
    keep = [ox, oy, obs_z]
    if locations is not None:
        loc = np.asarray(locations, dtype=np.float64)
        lx, ly = _f64(loc[:, 0]), _f64(loc[:, 1])
        keep += [lx, ly]
        npix = loc.shape[0]
        a.loc_x, a.loc_y = _ptr(lx), _ptr(ly)
        a.grid_ncol = 1
    else:
        if grid is None or npix is None:
            raise ValueError("give locations, or grid=(x0, y0, dx, dy, ncol) and npix")
        a.grid_x0, a.grid_y0, a.grid_dx, a.grid_dy = (float(v) for v in grid[:4])
        a.grid_ncol = int(grid[4])
        a.pix_begin = int(pix_begin)
    a.npix = int(npix)
    a.obs_x, a.obs_y, a.obs_z = _ptr(ox), _ptr(oy), _ptr(obs_z)
    a.nobs, a.ndays, a.n_obs, a.rm_dupl = nobs, ndays, int(n_obs), 1 if rm_dupl else 0
    cov_names = list(covariates)
    ncov = len(cov_names)
    cov_ptrs = (C.c_void_p * max(ncov, 1))()
    cov_strides = (C.c_int64 * max(ncov, 1))()
    rasters = (_CRaster * max(ncov, 1))()
    any_raster = False
    for i, name in enumerate(cov_names):
        cv = covariates[name]
        if isinstance(cv, Raster):      # sampled on the device at the prediction locations
            rasters[i], d = cv._c()
            keep.append(d)
            any_raster = True
            continue
        v = _f64(cv)
        if v.shape == (npix,):
            cov_strides[i] = 0
        elif v.shape == (ndays, npix):
            cov_strides[i] = npix
        else:
            raise ValueError(f"covariate {name!r} must have shape (npix,) or (ndays, npix), or be a Raster")
        cov_ptrs[i] = v.ctypes.data
        keep.append(v)
    a.ncov = ncov
    a.cov = C.cast(cov_ptrs, C.POINTER(C.c_void_p))
    a.cov_day_stride = C.cast(cov_strides, C.POINTER(C.c_int64))
    if any_raster:
        a.cov_raster = C.cast(rasters, C.POINTER(_CRaster))
    if cov_locations is not None:
        cl = np.asarray(cov_locations, dtype=np.float64)
        if cl.shape != (npix, 2):
            raise ValueError("cov_locations must be (npix, 2)")
        clx, cly = _f64(cl[:, 0]), _f64(cl[:, 1])
        keep += [clx, cly]
        a.cov_loc_x, a.cov_loc_y = _ptr(clx), _ptr(cly)
    fixes = (CovFix * max(len(cov_fix), 1))()
    for i, (t, o, delta) in enumerate(cov_fix):
        if t not in cov_names or o not in cov_names:
            raise ValueError(f"cov_fix: {t!r} / {o!r} are not covariates of this call")
        fixes[i] = CovFix(cov_names.index(t), cov_names.index(o), float(delta))
    if len(cov_fix):
        a.cov_fix = C.cast(fixes, C.POINTER(CovFix))
        a.ncov_fix = len(cov_fix)
    fmap = feature_map(model.feature_names, cov_names, int(n_obs))
    a.feat_map = fmap.ctypes.data_as(C.POINTER(C.c_int32))
    out = {}
    if want_mean:
        out["mean"] = np.empty((ndays, npix), dtype=np.float64)
        a.out_mean = _ptr(out["mean"])
        if centi_int16:
            out["mean_i16"] = np.empty((ndays, npix), dtype=np.int16)
            a.out_mean_i16 = _ptr(out["mean_i16"])
    probs = None
    if quantiles is not None:
        probs = _f64(quantiles)
        a.probs = probs.ctypes.data_as(C.POINTER(C.c_double))
        a.nprobs = len(probs)
        out["quantiles"] = np.empty((len(probs), ndays, npix), dtype=np.float64)
        a.out_q = _ptr(out["quantiles"])
        if centi_int16:
            out["iqr_i16"] = np.empty((ndays, npix), dtype=np.int16)
            a.out_iqr_i16 = _ptr(out["iqr_i16"])
    mask = None
    if pix_mask is not None:
        mask = np.ascontiguousarray(pix_mask, dtype=np.uint8)
        a.pix_mask = _ptr(mask)
    if devices is not None:
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        rc = lib.rfsi_predict_fused_multi(model.handle, C.byref(a), devs, len(devices))
    else:
        rc = lib.rfsi_predict_fused(model.handle, C.byref(a), device)
    if _lib.check(rc) == _lib.RFSI_W_TOO_FEW_OBS:
        warnings.warn(_lib.last_error())
    del keep, probs, mask
    return out
