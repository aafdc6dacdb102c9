#!/usr/bin/env python
"""Scale run: half of the 10^4 x 10^4 C5 raster (5 x 10^7 pixels) x 4 days through the HOST API in one
call -- covariates as per-day rasters sampled on the device, products as Int16 (what the mapping
scripts write) -- with a sample checked against the CPU oracle.
python tests/tools/scale_run.py [rows] [days] [ndevices] [trees]
With ndevices > 1 the same call is repeated with devices=0..ndevices-1 (rfsi_predict_fused_multi:
one band of rows per GPU inside ONE process) and must return the same bits."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle
import rfsi_b200 as rb
from rfsi_b200 import grow, synth

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 4
NDEV = int(sys.argv[3]) if len(sys.argv) > 3 else 1
TREES = int(sys.argv[4]) if len(sys.argv) > 4 else 500
case = synth.make_case("c5", ndays=20)
rng = np.random.default_rng(0)
# covariates: coarse daily rasters (1000 x 1000 cells of 10 x 10 pixels), Int16 tenths
R = {n: rb.Raster(np.round(10 * np.stack([synth.smooth_field(synth.lattice(1000, 1000, 5.0, 5.0, 10.0, 10.0), 100 + 17 * i + 1000 * d,
                                                             n_waves=16, length_scale=400.0).reshape(1000, 1000)[::-1]
                                          for d in range(D)])).astype(np.int16), 0.0, 10000.0, 10.0, 10.0, scale=0.1)
     for i, n in enumerate(case.cov_names)}
def np_extract(r, xy, d):
    c = np.minimum(np.floor((xy[:, 0] - r.x_ul) / r.dx), r.data.shape[2] - 1).astype(int)
    w = np.minimum(np.floor((r.y_ul - xy[:, 1]) / r.dy), r.data.shape[1] - 1).astype(int)
    return r.data[d][w, c].astype(np.float64) * r.scale + r.offset
case.cov_daily = {n: (lambda xy, d, n=n: np_extract(R[n], xy, d % D)) for n in R}
def gpu_no(l, o, z, n):
    a = rb.near_obs(l, np.column_stack([o, z]), zcol=2, n_obs=n).to_numpy()
    return a[:, :n], a[:, n:]
X, y = synth.training_table(case, gpu_no)
t0 = time.time(); flat = grow.grow_forest(X, y, case.feature_names, num_trees=TREES); print(f"grew {TREES} trees in {time.time()-t0:.1f}s", file=sys.stderr)
model = rb.RangerForest(flat); model.upload(0)
npix = rows * case.ncol
kw = dict(obs_xy=case.obs_xy, obs_z=case.obs_z[:D], n_obs=case.n_obs, grid=case.grid, npix=npix, covariates=R)
rb.predict_fused(model, **{**kw, "npix": 20 * case.ncol}, centi_int16=True)       # warm-up
t0 = time.perf_counter()
out = rb.predict_fused(model, centi_int16=True, **kw)
dt = time.perf_counter() - t0
print(f"{npix} pixels x {D} days = {npix*D:.3g} pixel-days in {dt:.2f} s = {npix*D/dt/1e6:.1f} M pixel-days/s (host API, rasters in, fp64 + Int16 out)")
print(f"  -> the full 10^8-pixel x 365-day grid: {1e8*365/(npix*D/dt)/60:.1f} min on one B200, {1e8*365/(npix*D/dt)/60/7.8:.1f} min on 8")
sel = np.sort(rng.choice(npix, 2000, replace=False))
loc = case.locations(0, npix)[sel]
ref = oracle.predict_fused(flat.as_dict(), case.feature_names, loc_xy=loc, obs_xy=case.obs_xy, obs_z=case.obs_z[:D], n_obs=case.n_obs,
                           covariates={n: np.stack([np_extract(R[n], loc, d) for d in range(D)]) for n in R}, kdtree=True, threads=16)
assert np.array_equal(out["mean"][:, sel], ref["mean"]), "sample differs from the oracle"
assert np.array_equal(out["mean_i16"][:, sel].ravel(), oracle.centi_int16(ref["mean"].ravel()))
print("  sample of 2000 pixels x all days: bit-equal to the CPU oracle (fp64 mean and Int16 product)")
if NDEV > 1:
    devs = list(range(min(NDEV, rb.device_count()))) if rb.device_count() > 1 else [0] * NDEV
    for d in set(devs):
        model.upload(d)
    rb.predict_fused(model, **{**kw, "npix": 20 * case.ncol * len(devs)}, centi_int16=True, devices=devs)   # warm-up: contexts, pools
    t0 = time.perf_counter()
    multi = rb.predict_fused(model, centi_int16=True, devices=devs, **kw)
    dtm = time.perf_counter() - t0
    assert all(np.array_equal(out[k], multi[k]) for k in out), "banded result differs from the single-device call"
    print(f"devices={devs} in ONE process: {dtm:.2f} s = {npix*D/dtm/1e6:.1f} M pixel-days/s ({dt/dtm:.2f}x the single-device call), bit-equal")
