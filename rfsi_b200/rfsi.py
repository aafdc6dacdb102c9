"""rfsi() / pred_rfsi(): the wrappers the reference documents in README.md:93-143.

rfsi   = near.obs on the training stations (locations == observations, per day) + cbind +
         ranger(); here the grower in csrc/grow.cpp stands in for ranger() -- training is not
         on the accelerated path.
pred_rfsi = near.obs + cbind + predict for new locations, in one fused engine call.

Inputs are data frames (the README's `data.frame` branch: data.staid.x.y.z names the station
id, x, y and time columns; time None/NA = purely spatial).  sf / SpatVector / SpatRaster
objects have no Python counterpart here; pass their attribute table + coordinates.
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
    if output_format != "data.frame":
        raise NotImplementedError("pred_rfsi: only output.format = 'data.frame' exists without terra/sf")
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
