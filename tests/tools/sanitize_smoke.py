#!/usr/bin/env python
"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck):
   compute-sanitizer --tool memcheck python tests/tools/sanitize_smoke.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import rfsi_b200 as rb  # noqa: E402
from rfsi_b200 import grow, synth  # noqa: E402

rng = np.random.default_rng(0)
# brute kNN (TMA tiles), cells kNN (+ location sort), FINAL and multi-day forms
loc = synth.lattice(40, 37)
obs = loc[rng.choice(len(loc), 90, replace=False)]
z = rng.normal(size=90)
rb.near_obs(loc, np.column_stack([obs, z]), zcol=2, n_obs=7)
big = rng.uniform(0, 100, size=(1500, 2))
rb.near_obs(rng.uniform(0, 100, size=(999, 2)), np.column_stack([big, rng.normal(size=1500)]), zcol=2, n_obs=5)
zt = rng.normal(size=(3, 90)); zt[1, ::2] = np.nan; zt[2, 5:] = np.nan
rb.near_obs_days(obs, zt, 6)
# forest (every image), quantiles, terminal nodes
X = rng.normal(size=(700, 9)); y = X[:, 0] + rng.normal(size=700)
flat = grow.grow_forest(X, y, [f"f{i}" for i in range(9)], num_trees=21, quantreg=True)
import os
for compact in ("3", "2", "1", "0"):
    os.environ["RFSI_B200_COMPACT"] = compact
    m = rb.RangerForest(flat)
    Xp = rng.normal(size=(1001, 9))
    rb.predict(m, Xp); rb.predict(m, Xp, type="quantiles"); rb.predict(m, Xp, type="terminalNodes")
    m.close()
del os.environ["RFSI_B200_COMPACT"]       # default image (sector) with its rank caches
# fused: raster grid + explicit, fallback (many missing), QRF + Int16, raster covariate, mask
case = synth.make_case("c4", ndays=3)
case.obs_z[1, rng.choice(87, 60, replace=False)] = np.nan
import oracle  # training table only
Xt, yt = synth.training_table(case, lambda l, o, zz, n: oracle.near_obs(l, o, zz, n)[:2])
fm = rb.RangerForest(grow.grow_forest(Xt, yt, case.feature_names, num_trees=17, quantreg=True))
npix = 31 * case.ncol
locs = case.locations(0, npix)
cov = case.covariates(locs)
cov["imerg"] = rb.Raster(rng.integers(0, 50, size=(3, 30, 40)).astype(np.uint16), 0.0, 300.0, 10.0, 10.0, scale=0.1)
mask = (rng.uniform(size=npix) < 0.6).astype(np.uint8)
rb.predict_fused(fm, obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, grid=case.grid, npix=npix,
                 covariates=cov, quantiles=[.25, .5, .75], centi_int16=True, pix_mask=mask)
rb.predict_fused(fm, obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, locations=locs[::3],
                 covariates={k: (v if isinstance(v, rb.Raster) else v[:, ::3]) for k, v in cov.items()})
print("sanitize_smoke: done")
