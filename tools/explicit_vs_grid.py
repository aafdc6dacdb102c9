#!/usr/bin/env python
"""Fused prediction with the raster given as explicit coordinates (what R passes) vs the grid form."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import rfsi_b200 as rb
from rfsi_b200 import grow, synth

def gpu_no(l, o, z, n):
    a = rb.near_obs(l, np.column_stack([o, z]), zcol=2, n_obs=n).to_numpy()
    return a[:, :n], a[:, n:]

case = synth.make_case("c5", ndays=6)
X, y = synth.training_table(case, gpu_no)
model = rb.RangerForest(grow.grow_forest(X, y, case.feature_names, num_trees=100))
rows, D = 60, 4
npix = rows * case.ncol
loc = case.locations(0, npix)
cov = {n: np.stack([case.cov_daily[n](loc, d) for d in range(D)]) for n in case.cov_names}
kw = dict(obs_xy=case.obs_xy, obs_z=case.obs_z[:D], n_obs=case.n_obs, covariates=cov)
def best(fn):
    fn(); ts = []
    for _ in range(3):
        t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
    return min(ts), r
tg, rg = best(lambda: rb.predict_fused(model, grid=case.grid, npix=npix, **kw))
te, re_ = best(lambda: rb.predict_fused(model, locations=loc, **kw))
perm = np.random.default_rng(0).permutation(npix)
loc_s, cov_s = np.ascontiguousarray(loc[perm]), {k: np.ascontiguousarray(v[:, perm]) for k, v in cov.items()}
ts, rs = best(lambda: rb.predict_fused(model, locations=loc_s, covariates=cov_s, obs_xy=case.obs_xy,
                                       obs_z=case.obs_z[:D], n_obs=case.n_obs))
assert np.array_equal(rg["mean"], re_["mean"]) and np.array_equal(rg["mean"][:, perm], rs["mean"])
print(f"grid form {npix*D/tg/1e6:.1f} M pixel-days/s | explicit row-major coords {npix*D/te/1e6:.1f} M | shuffled coords {npix*D/ts/1e6:.1f} M")
