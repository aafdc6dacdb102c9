"""Synthetic stations, rasters and forests of the shapes BASELINE.json names.

The reference ships no trained forest (.MISSING_LARGE_BLOBS) and no network is
available, so benchmarks and tests build their own inputs: NumPy
default_rng(seed) (PCG64), fp64.  Forests are *grown* here only to have realistic
tree shapes to traverse; training is not part of the accelerated path.

The near.obs implementation used to build training tables is passed in by the caller
(the GPU near_obs in bench.py, the CPU oracle in CPU-only tests).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

from .forest import FlatForest, from_sklearn


def smooth_field(xy: np.ndarray, seed: int, n_waves: int = 64, length_scale: float = 50.0) -> np.ndarray:
    """Zero-mean, unit-variance-ish smooth random field: a sum of random cosines."""
    rng = np.random.default_rng(seed)
    k = rng.normal(size=(n_waves, 2)) / length_scale
    ph = rng.uniform(0, 2 * np.pi, size=n_waves)
    out = np.zeros(xy.shape[0])
    step = 1 << 18
    for a in range(0, xy.shape[0], step):
        out[a:a + step] = np.sqrt(2.0 / n_waves) * np.cos(xy[a:a + step] @ k.T + ph).sum(axis=1)
    return out


def lattice(ncol: int, nrow: int, x0: float = 0.5, y0: float = 0.5, dx: float = 1.0, dy: float = 1.0) -> np.ndarray:
    """Cell centres in row-major raster order: cell p -> (x0 + (p % ncol)*dx, y0 + (p // ncol)*dy)."""
    c = np.arange(ncol, dtype=np.float64)
    r = np.arange(nrow, dtype=np.float64)
    x = x0 + c * dx
    y = y0 + r * dy
    return np.stack([np.tile(x, nrow), np.repeat(y, ncol)], axis=1)


@dataclass
class Case:
    name: str
    grid: tuple                    # (x0, y0, dx, dy, ncol), row-major raster
    nrow: int
    obs_xy: np.ndarray             # (S, 2)
    obs_z: np.ndarray              # (ndays, S), NaN = missing that day
    n_obs: int
    cov_names: list
    cov_static: dict = field(default_factory=dict)   # name -> callable(xy) -> values
    cov_daily: dict = field(default_factory=dict)    # name -> callable(xy, day) -> values
    feature_names: list = field(default_factory=list)

    @property
    def ncol(self) -> int:
        return int(self.grid[4])

    @property
    def npix(self) -> int:
        return self.ncol * self.nrow

    def locations(self, begin: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.npix - begin if count is None else count
        p = np.arange(begin, begin + count)
        x0, y0, dx, dy, ncol = self.grid
        return np.stack([x0 + (p % ncol).astype(np.float64) * dx, y0 + (p // ncol).astype(np.float64) * dy], axis=1)

    def covariates(self, xy: np.ndarray, ndays: Optional[int] = None) -> dict:
        ndays = self.obs_z.shape[0] if ndays is None else ndays
        out = {}
        for n in self.cov_names:
            if n in self.cov_static:
                out[n] = self.cov_static[n](xy)
            else:
                out[n] = np.stack([self.cov_daily[n](xy, d) for d in range(ndays)], axis=0)
        return out


def rfsi_feature_names(n_obs: int, cov_names=()) -> list:
    """Covariates, then dist1..distN, then obs1..obsN (cbind(df, nearest_obs), simulation.R:199)."""
    return list(cov_names) + [f"dist{j + 1}" for j in range(n_obs)] + [f"obs{j + 1}" for j in range(n_obs)]


def make_case(name: str, seed: int = 42, ndays: Optional[int] = None) -> Case:
    """The BASELINE.json / SURVEY.md 8(d) shapes.

    c1        simulation.R as named: 100x100 lattice, 200 stations drawn from the cell
              centres (self matches + exact ties), n = 6
    c1s       simulation.R as scripted: 500x500, 1000 stations from the centres, n = 25
    c2        sim_500.R: 500x500, 500 stations from the centres, n = 40
    c3        temp_croatia shape: 490x487 cells of 1 km, 157 stations, 3-6 % missing per
              day, n = 10, 8 covariates (3 static, 5 daily)
    c4        prcp_catalonia shape: 380x280, 87 stations, 1-5 missing per day, n = 7,
              3 daily covariates
    c5        scale-up: 10^4 x 10^4 cells, 10^4 stations (continuous coordinates: no
              ties), 2 % missing per day, n = 20, 3 daily covariates
    """
    rng = np.random.default_rng(seed)

    def stations_from_lattice(ncol, nrow, S):
        sel = rng.choice(ncol * nrow, S, replace=False)
        return lattice(ncol, nrow)[sel]

    def z_table(xy, D, ls, frac_missing=(0.0, 0.0), seasonal=False):
        base = 20.0 + np.sqrt(10.0) * smooth_field(xy, seed + 1, length_scale=ls)
        z = np.empty((D, xy.shape[0]))
        for d in range(D):
            z[d] = base + rng.normal(0.0, np.sqrt(2.5), size=xy.shape[0])
            if seasonal:
                z[d] += 8.0 * np.sin(2 * np.pi * d / 365.0)
            lo, hi = frac_missing
            if hi > 0:
                m = int(round(rng.uniform(lo, hi) * xy.shape[0]))
                if m:
                    z[d, rng.choice(xy.shape[0], m, replace=False)] = np.nan
        return z

    def daily_cov(i, ls):
        return lambda xy, d: smooth_field(xy, seed + 100 + 17 * i + 1000 * d, n_waves=16, length_scale=ls)

    def static_cov(i, ls):
        return lambda xy: smooth_field(xy, seed + 50 + i, n_waves=16, length_scale=ls)

    if name == "c1":
        xy = stations_from_lattice(100, 100, 200)
        return Case(name, (0.5, 0.5, 1.0, 1.0, 100), 100, xy, z_table(xy, ndays or 1, 50.0), 6, [],
                    feature_names=rfsi_feature_names(6))
    if name == "c1s":
        xy = stations_from_lattice(500, 500, 1000)
        return Case(name, (0.5, 0.5, 1.0, 1.0, 500), 500, xy, z_table(xy, ndays or 1, 50.0), 25, [],
                    feature_names=rfsi_feature_names(25))
    if name == "c2":
        xy = stations_from_lattice(500, 500, 500)
        return Case(name, (0.5, 0.5, 1.0, 1.0, 500), 500, xy, z_table(xy, ndays or 1, 100.0), 40, [],
                    feature_names=rfsi_feature_names(40))
    if name == "c3":
        x0, y0, dx, dy, ncol, nrow = 359641.0 + 500.0, 5156149.0 - 500.0, 1000.0, -1000.0, 490, 487
        xy = np.stack([rng.uniform(387331.0, 842443.0, 157), rng.uniform(4694138.0, 5137646.0, 157)], axis=1)
        covs = [f"cov{i + 1}" for i in range(8)]
        c = Case(name, (x0, y0, dx, dy, ncol), nrow, xy,
                 z_table(xy, ndays or 4, 80000.0, (0.03, 0.06), seasonal=True), 10, covs,
                 feature_names=rfsi_feature_names(10, covs))
        for i in range(3):
            c.cov_static[covs[i]] = static_cov(i, 60000.0)
        for i in range(3, 8):
            c.cov_daily[covs[i]] = daily_cov(i, 90000.0)
        return c
    if name == "c4":
        x0, y0, dx, dy, ncol, nrow = 0.5, 0.5, 1.0, 1.0, 380, 280
        xy = np.stack([rng.uniform(0, 380, 87), rng.uniform(0, 280, 87)], axis=1)
        covs = ["imerg", "tmax", "tmin"]
        c = Case(name, (x0, y0, dx, dy, ncol), nrow, xy,
                 z_table(xy, ndays or 4, 60.0, (1.0 / 87, 5.0 / 87)), 7, covs,
                 feature_names=rfsi_feature_names(7, covs))
        for i, n in enumerate(covs):
            c.cov_daily[n] = daily_cov(i, 70.0)
        return c
    if name == "c5":
        ncol = nrow = 10000
        S = 10000
        xy = np.stack([rng.uniform(0, ncol, S), rng.uniform(0, nrow, S)], axis=1)
        covs = ["cov1", "cov2", "cov3"]
        c = Case(name, (0.5, 0.5, 1.0, 1.0, ncol), nrow, xy,
                 z_table(xy, ndays or 8, 500.0, (0.02, 0.02), seasonal=True), 20, covs,
                 feature_names=rfsi_feature_names(20, covs))
        for i, n in enumerate(covs):
            c.cov_daily[n] = daily_cov(i, 400.0)
        return c
    raise ValueError(f"unknown case {name!r}")


def training_table(case: Case, near_obs_fn: Callable, max_rows: Optional[int] = None, seed: int = 7):
    """Per-day near.obs with locations == observations (temp_croatia/1_models.R:1034-1047):
    every station's own observation is excluded.  Returns (X[n, F], y[n])."""
    rng = np.random.default_rng(seed)
    Xs, ys = [], []
    for d in range(case.obs_z.shape[0]):
        valid = ~np.isnan(case.obs_z[d])
        xy = case.obs_xy[valid]
        z = case.obs_z[d][valid]
        dist, obs = near_obs_fn(xy, xy, z, case.n_obs)
        cov = {}
        for n in case.cov_names:
            cov[n] = case.cov_static[n](xy) if n in case.cov_static else case.cov_daily[n](xy, d)
        cols = []
        for n in case.feature_names:
            if n.startswith("dist"):
                cols.append(dist[:, int(n[4:]) - 1])
            elif n.startswith("obs"):
                cols.append(obs[:, int(n[3:]) - 1])
            else:
                cols.append(cov[n])
        X = np.stack(cols, axis=1)
        ok = ~np.isnan(X).any(axis=1)            # complete.cases, 1_models.R:1003
        Xs.append(X[ok]); ys.append(z[ok])
    X = np.concatenate(Xs); y = np.concatenate(ys)
    if max_rows is not None and X.shape[0] > max_rows:
        sel = rng.choice(X.shape[0], max_rows, replace=False)
        X, y = X[sel], y[sel]
    return X, y


def grow_forest_sklearn(X, y, feature_names, num_trees=250, min_node_size=5, mtry=None, quantreg=False,
                        seed=42, n_jobs=-1) -> FlatForest:
    """ranger(..., num.trees, mtry, min.node.size) stand-in: scikit-learn grows the trees."""
    from sklearn.ensemble import RandomForestRegressor
    F = X.shape[1]
    mtry = mtry or max(1, int(np.floor(np.sqrt(F))))
    rf = RandomForestRegressor(n_estimators=num_trees, max_features=mtry, min_samples_split=max(2, min_node_size),
                               bootstrap=True, random_state=seed, n_jobs=n_jobs)
    rf.fit(X, y)
    return from_sklearn(rf, feature_names, X_train=X, y_train=y, quantreg=quantreg, seed=seed)
