"""rfsi() / pred_rfsi(): the wrappers the reference documents in README.md:93-143.

rfsi   = near.obs on the training stations (locations == observations, per day) + cbind +
         ranger(); here the grower in csrc/grow.cpp stands in for ranger() -- training is not
         on the accelerated path.
pred_rfsi = near.obs + cbind + predict for new locations, in one fused engine call.

Inputs are data frames (the README's `data.frame` branch: data.staid.x.y.z names the station
id, x, y and time columns; time None/NA = purely spatial).  sf / SpatVector objects have no
Python counterpart here: pass their attribute table + coordinates.  A SpatRaster of covariates
(README.md:121-125, `newdata <- terra::rast(meuse.grid)`) is a dict of api.Raster layers on one
grid (e.g. read with geotiff.read_geotiff); with output_format="SpatRaster" the prediction
comes back as layers on that grid (README.md:133,150-153).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import pandas as pd

from . import api, grow
from .forest import RangerForest


def _parse_formula(formula: str):
    lhs, rhs = formula.split("~")
    terms = [t.strip() for t in rhs.split("+") if t.strip() and t.strip() != "1"]
    return lhs.strip(), terms


def _col(df: pd.DataFrame, c):
    return df.columns[c] if isinstance(c, (int, np.integer)) else c


@dataclass
class RfsiModel:
    forest: RangerForest
    formula: str
    response: str
    covariates: list
    n_obs: int
    num_trees: int
    quantreg: bool

    @property
    def feature_names(self):
        return self.forest.feature_names


def rfsi(formula: str, data: pd.DataFrame, data_staid_x_y_z: Sequence = (0, 1, 2, None), n_obs: int = 5,
         zero_tol: float = 0, cpus: int = 0, progress: bool = False, num_trees: int = 250, mtry: Optional[int] = None,
         min_node_size: int = 5, quantreg: bool = False, seed: int = 42, device: int = 0) -> RfsiModel:
    """README.md:93-110.  Returns the fitted model (forest resident in the engine)."""
    if zero_tol != 0:
        raise NotImplementedError("rfsi: only zero.tol = 0 is supported")
    response, covs = _parse_formula(formula)
    sid, xc, yc, tc = [None if c is None else _col(data, c) for c in data_staid_x_y_z]
    parts = []
    if tc is None:
        g = data[~data[response].isna()]
        no = api.near_obs(g[[xc, yc]], g[[xc, yc, response]], zcol=2, n_obs=n_obs, device=device)
        parts.append(pd.concat([g[[response] + covs].reset_index(drop=True), no], axis=1))
    else:
        # every day in ONE engine call (rfsi_near_obs_days_f64): stations x days table, NaN = silent
        stations = data[[sid, xc, yc]].drop_duplicates(sid).reset_index(drop=True)
        pos = {s: i for i, s in enumerate(stations[sid])}
        days = sorted(data[tc].unique())
        di = np.searchsorted(np.asarray(days), data[tc].to_numpy())
        si = data[sid].map(pos).to_numpy()
        z = np.full((len(days), len(stations)), np.nan)
        z[di, si] = data[response].to_numpy(dtype=np.float64)
        dist, obs = api.near_obs_days(stations[[xc, yc]].to_numpy(dtype=np.float64), z, n_obs, device=device)
        keep = ~data[response].isna().to_numpy()
        tab = data.loc[keep, [response] + covs].reset_index(drop=True)
        for j in range(n_obs):
            tab[f"dist{j + 1}"] = dist[di[keep], si[keep], j]
        for j in range(n_obs):
            tab[f"obs{j + 1}"] = obs[di[keep], si[keep], j]
        parts.append(tab)
    table = pd.concat(parts, ignore_index=True).dropna()          # complete.cases (1_models.R:1003)
    names = covs + [f"dist{j + 1}" for j in range(n_obs)] + [f"obs{j + 1}" for j in range(n_obs)]
    flat = grow.grow_forest(table[names].to_numpy(), table[response].to_numpy(), names, num_trees=num_trees,
                            mtry=mtry, min_node_size=min_node_size, seed=seed, num_threads=cpus, quantreg=quantreg)
    return RfsiModel(RangerForest(flat), formula, response, covs, n_obs, num_trees, quantreg)


def pred_rfsi(model: RfsiModel, data: pd.DataFrame, obs_col, data_staid_x_y_z: Sequence, newdata: pd.DataFrame,
              newdata_staid_x_y_z: Sequence, output_format: str = "data.frame", zero_tol: float = 0, cpus: int = 1,
              progress: bool = False, soil3d: bool = False, type: str = "response",
              quantiles: Sequence[float] = (0.1, 0.5, 0.9), device: int = 0) -> pd.DataFrame:
    """README.md:127-143.  Returns newdata's id/x/y(/time) + `pred` (+ `quantile..p` columns)."""
    if zero_tol != 0 or soil3d:
        raise NotImplementedError("pred_rfsi: zero.tol = 0 and soil3d = FALSE only")
    if isinstance(newdata, dict) and all(isinstance(v, api.Raster) for v in newdata.values()):
        if output_format != "SpatRaster":
            raise NotImplementedError("pred_rfsi: a raster newdata gives output.format = 'SpatRaster'")
        return _pred_rfsi_raster(model, data, obs_col, data_staid_x_y_z, newdata, type, quantiles, device)
    if output_format != "data.frame":
        raise NotImplementedError("pred_rfsi: output.format = 'data.frame' (or a raster newdata with 'SpatRaster')")
    if type not in ("response", "quantiles"):
        raise ValueError("type must be 'response' or 'quantiles'")
    if type == "quantiles" and not model.quantreg:
        raise ValueError("Error: Set quantreg=TRUE in rfsi(...) for quantile prediction.")
    sid, xc, yc, tc = [None if c is None else _col(data, c) for c in data_staid_x_y_z]
    nsid, nxc, nyc, ntc = [None if c is None else _col(newdata, c) for c in newdata_staid_x_y_z]
    obs_col = _col(data, obs_col)
    stations = data[[sid, xc, yc]].drop_duplicates(sid).reset_index(drop=True)
    pos = {s: i for i, s in enumerate(stations[sid])}
    days = [None] if tc is None else sorted(data[tc].unique())
    z = np.full((len(days), len(stations)), np.nan)
    di = np.zeros(len(data), dtype=np.int64) if tc is None else np.searchsorted(np.asarray(days), data[tc].to_numpy())
    z[di, data[sid].map(pos).to_numpy()] = data[obs_col].to_numpy(dtype=np.float64)
    probs = list(quantiles) if type == "quantiles" else None
    out_parts = []
    new_groups = [(None, newdata)] if ntc is None else list(newdata.groupby(ntc, sort=True))
    for key, g in new_groups:
        d = 0 if key is None or tc is None else days.index(key)
        cov = {c: g[c].to_numpy(dtype=np.float64) for c in model.covariates}
        res = api.predict_fused(model.forest, obs_xy=stations[[xc, yc]].to_numpy(dtype=np.float64), obs_z=z[d:d + 1],
                                n_obs=model.n_obs, locations=g[[nxc, nyc]].to_numpy(dtype=np.float64),
                                covariates=cov, quantiles=probs, device=device)
        keep = [c for c in (nsid, nxc, nyc, ntc) if c is not None]
        part = g[keep].reset_index(drop=True)
        part["pred"] = res["mean"][0]
        if probs is not None:
            for j, p in enumerate(probs):
                part[f"quantile..{p}"] = res["quantiles"][j, 0]
        out_parts.append(part)
    return pd.concat(out_parts, ignore_index=True)


@dataclass
class RasterPrediction:
    """pred.rfsi(..., output.format = "SpatRaster"): layers `pred` (+ `quantile..p`) on newdata's grid,
    NaN where a covariate layer is NA (README.md:150-153)."""
    layers: dict
    x_ul: float
    y_ul: float
    dx: float
    dy: float

    def __getitem__(self, name):
        return self.layers[name]

    def names(self):
        return list(self.layers)


def raster_cells(r: "api.Raster"):
    """Cell centres of a north-up raster in row-major order (xyFromCell) and its (nrow, ncol)."""
    d = r.data if r.data.ndim == 2 else r.data[0]
    nrow, ncol = d.shape
    x = r.x_ul + (np.arange(ncol) + 0.5) * r.dx
    y = r.y_ul - (np.arange(nrow) + 0.5) * r.dy
    return np.column_stack([np.tile(x, nrow), np.repeat(y, ncol)]), (nrow, ncol)


def raster_valid(r: "api.Raster") -> np.ndarray:
    """!is.na(values(r)) of the first layer, row-major."""
    d = (r.data if r.data.ndim == 2 else r.data[0]).astype(np.float64)
    ok = ~np.isnan(d)
    if not np.isnan(r.nodata):
        ok &= d != r.nodata
    return ok.ravel()


def _pred_rfsi_raster(model: RfsiModel, data: pd.DataFrame, obs_col, data_staid_x_y_z, newdata: dict, type: str,
                      quantiles, device: int) -> RasterPrediction:
    if type not in ("response", "quantiles"):
        raise ValueError("type must be 'response' or 'quantiles'")
    if type == "quantiles" and not model.quantreg:
        raise ValueError("Error: Set quantreg=TRUE in rfsi(...) for quantile prediction.")
    missing = [c for c in model.covariates if c not in newdata]
    if missing:
        raise ValueError(f"pred_rfsi: newdata has no layer for {missing}")
    sid, xc, yc, tc = [None if c is None else _col(data, c) for c in data_staid_x_y_z]
    if tc is not None and data[tc].nunique() > 1:
        raise NotImplementedError("pred_rfsi: a raster newdata is purely spatial (one time step)")
    obs_col = _col(data, obs_col)
    layers = {c: newdata[c] for c in model.covariates}
    first = next(iter(newdata.values()))
    loc, (nrow, ncol) = raster_cells(first)
    for name, r in layers.items():
        d = r.data if r.data.ndim == 2 else r.data[0]
        if (d.shape, r.x_ul, r.y_ul, r.dx, r.dy) != ((nrow, ncol), first.x_ul, first.y_ul, first.dx, first.dy):
            raise ValueError(f"pred_rfsi: layer {name!r} is not on the grid of the first layer")
    valid = np.ones(nrow * ncol, dtype=bool)
    for r in layers.values():
        valid &= raster_valid(r)
    if not layers:
        valid &= raster_valid(first)
    probs = list(quantiles) if type == "quantiles" else None
    z = data[obs_col].to_numpy(dtype=np.float64)[None, :]
    res = api.predict_fused(model.forest, obs_xy=data[[xc, yc]].to_numpy(dtype=np.float64), obs_z=z, n_obs=model.n_obs,
                            locations=loc, covariates=layers, quantiles=probs, pix_mask=valid.astype(np.uint8),
                            device=device)
    out = {"pred": res["mean"][0].reshape(nrow, ncol)}
    if probs is not None:
        for j, p in enumerate(probs):
            out[f"quantile..{p}"] = res["quantiles"][j, 0].reshape(nrow, ncol)
    return RasterPrediction(out, first.x_ul, first.y_ul, first.dx, first.dy)
