"""Mint the golden vectors under tests/golden/ from the CPU oracle.

The reference holds no golden vectors for this path, so these are the
repo's own: inputs from seeded generators, expected outputs from oracle/rfsi_oracle.c,
cross-checked at mint time against the independent NumPy and kd-tree restatements.
Run from the repo root:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
from oracle import rfsi_oracle_np as onp  # noqa: E402
from rfsi_b200 import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


def knn_lattice():
    """C1-named: 100x100 lattice, 200 stations from the centres (ties + self matches), n=6."""
    case = synth.make_case("c1")
    loc = case.locations()
    z = case.obs_z[0]
    d, o, i, _ = oracle.near_obs(loc, case.obs_xy, z, 6)
    d2, o2, i2, _ = oracle.near_obs(loc, case.obs_xy, z, 6, kdtree=True, nthreads=2)
    assert np.array_equal(d, d2) and np.array_equal(i, i2) and np.array_equal(o, o2)
    sub = np.arange(0, loc.shape[0], 7)
    d3, o3, i3 = onp.near_obs(loc[sub], case.obs_xy, z, 6)
    assert np.array_equal(d[sub], d3) and np.array_equal(i[sub], i3)
    np.savez_compressed(OUT / "knn_c1_lattice.npz", grid=np.array(case.grid), nrow=case.nrow, obs_xy=case.obs_xy,
                        obs_z=z, n_obs=6, dist=d, obs=o, idx=i)


def knn_realgeom():
    """C3 geometry (UTM-sized coordinates, 1 km cells), first 8 raster rows, n=10."""
    case = synth.make_case("c3", ndays=1)
    loc = case.locations(0, 8 * case.ncol)
    valid = ~np.isnan(case.obs_z[0])
    d, o, i, _ = oracle.near_obs(loc, case.obs_xy[valid], case.obs_z[0][valid], 10)
    d2, _, i2, _ = oracle.near_obs(loc, case.obs_xy[valid], case.obs_z[0][valid], 10, kdtree=True, nthreads=2)
    assert np.array_equal(d, d2) and np.array_equal(i, i2)
    np.savez_compressed(OUT / "knn_c3_utm.npz", grid=np.array(case.grid), npix=loc.shape[0],
                        obs_xy=case.obs_xy[valid], obs_z=case.obs_z[0][valid], n_obs=10, dist=d, obs=o, idx=i)


def forest_small():
    """sklearn-grown 40-tree quantile forest on C1 features; mean, terminal nodes, quantiles."""
    case = synth.make_case("c1", ndays=3)
    X, y = synth.training_table(case, lambda l, o, z, n: oracle.near_obs(l, o, z, n)[:2])
    flat = synth.grow_forest_sklearn(X, y, case.feature_names, num_trees=40, quantreg=True, n_jobs=1)
    loc = case.locations()[::13]
    d, o, _, _ = oracle.near_obs(loc, case.obs_xy, case.obs_z[0], 6)
    Xp = np.concatenate([d, o], axis=1)
    fo = flat.as_dict()
    mean = oracle.forest_predict(fo, Xp)
    term = oracle.forest_terminal_nodes(fo, Xp)
    probs = np.array([0.25, 0.5, 0.75])
    q = oracle.forest_quantiles(fo, Xp, probs)
    assert np.array_equal(mean[:50], onp.forest_predict(fo, Xp[:50]))
    assert np.array_equal(q[:50], onp.forest_quantiles(fo, Xp[:50], probs))
    assert np.array_equal(mean, oracle.forest_predict(fo, Xp, threads=2))
    np.savez_compressed(OUT / "forest_c1_40trees.npz", X=Xp, probs=probs, mean=mean, term=term, q=q,
                        feature_names=np.array(flat.feature_names),
                        **{k: v for k, v in fo.items() if v is not None})


def knn_croatia_real():
    """Real station geometry: temp_croatia/temp_data/stfdf.rda (157 stations, UTM33 metres) on
    day 5 (the first mapped slice, 2_predictions.R:65), stations with TEMP observed that day
    (2_predictions.R:122-125), against 8 rows of the dem.tif grid (490 columns, 1 km cells,
    tie-point 359641 / 5156149), n.obs = 10 (2_predictions.R:159).  Inputs come from the
    reference's data file through rfsi_b200/rda.py; expected outputs from the oracle."""
    from rfsi_b200 import rda
    ref = Path("/root/reference/temp_croatia/temp_data/stfdf.rda")
    st = rda.read_rda(ref)["stfdf"]
    xy = np.asarray(st.attrs["sp"].attrs["coords"], dtype=np.float64)
    temp = np.asarray(st.attrs["data"]["TEMP"], dtype=np.float64).reshape(366, 157)   # space index runs fastest
    day = 4
    ok = ~np.isnan(temp[day])
    grid = (359641.0 + 500.0, 5156149.0 - 500.0, 1000.0, -1000.0, 490)
    p = np.arange(200 * 490, 208 * 490)
    loc = np.stack([grid[0] + (p % 490) * grid[2], grid[1] + (p // 490) * grid[3]], axis=1)
    d, o, i, _ = oracle.near_obs(loc, xy[ok], temp[day][ok], 10)
    d2, _, i2, _ = oracle.near_obs(loc, xy[ok], temp[day][ok], 10, kdtree=True, nthreads=2)
    assert np.array_equal(d, d2) and np.array_equal(i, i2)
    np.savez_compressed(OUT / "knn_croatia_real.npz", grid=np.array(grid), pix_begin=200 * 490, npix=len(p),
                        obs_xy=xy, temp_days=temp[:8], day=day, n_obs=10, dist=d, obs=o, idx=i)


def croatia_llocv_inputs():
    """The real inputs of the Croatian case study (temp_croatia/temp_data/stfdf.rda + folds.rda):
    157 stations x 366 days of TEMP and the covariates of 1_models.R:73, and the 10 station
    folds of the nested LLOCV (1_models.R:1094-1233).  Used by the end-to-end accuracy test,
    whose anchor is the reference's own stored result: RMSE 1.3988, CCC 0.9827
    (1_models.R:1245-1249; temp_data/rfsi_10f_{obs,pred}.rda)."""
    from rfsi_b200 import rda
    base = Path("/root/reference/temp_croatia/temp_data")
    st = rda.read_rda(base / "stfdf.rda")["stfdf"]
    sp = st.attrs["sp"]
    spd, d = sp.attrs["data"], st.attrs["data"]
    daily = {n: np.asarray(d[n], dtype=np.float64).reshape(366, 157) for n in ("TEMP", "INSOL", "ctd", "MODIS.LST")}
    static = {n: np.asarray(spd[n], dtype=np.float64) for n in ("HRdem", "HRdsea", "Lat", "Lon", "HRtwi")}
    folds = rda.read_rda(base / "folds.rda")["strat"]["obs.fold.list"]
    fold_of = np.zeros(157, dtype=np.int8)
    for k, f in enumerate(folds):
        fold_of[np.asarray(f) - 1] = k + 1
    obs = np.asarray(list(rda.read_rda(base / "rfsi_10f_obs.rda").values())[0], dtype=np.float64)
    pred = np.asarray(list(rda.read_rda(base / "rfsi_10f_pred.rda").values())[0], dtype=np.float64)
    ok = ~np.isnan(obs)
    np.savez_compressed(OUT / "croatia_llocv.npz", xy=np.asarray(sp.attrs["coords"], dtype=np.float64),
                        TEMP=daily["TEMP"], INSOL=daily["INSOL"].astype(np.float32), ctd=daily["ctd"].astype(np.float32),
                        MODIS_LST=daily["MODIS.LST"].astype(np.float32),
                        **{k: v.astype(np.float32) for k, v in static.items()}, fold=fold_of,
                        ref_rmse=np.sqrt(np.mean((obs[ok] - pred[ok]) ** 2)), ref_n=int(ok.sum()))


def croatia_idw_cv():
    """A reference-produced output vector that depends on near.obs's substance: the stored
    predictions of the nested 10-fold IDW cross-validation on the Croatian temperatures
    (temp_croatia/1_models.R:405-536 -> temp_data/IDW_10f_pred.rda, computed in R by gstat with
    `nmax = n, idp = p` on the UTM33 station coordinates).  An IDW prediction is a function of
    exactly what near.obs returns -- the n nearest stations with an observation that day, their
    distances and their observed values: pred = sum(z_i d_i^-p) / sum(d_i^-p).  The per-fold
    (n, p) the script's tuning loop arrived at are not recorded in the script; they are
    recovered here by searching its tuning grid (p = 1.0 .. 3.5 step 0.1, n = 2 .. 40): for every
    fold exactly one pair reproduces all of the fold's stored predictions to < 3e-14 (the next
    best pair is off by > 1e-3).  Inputs (coordinates, TEMP, folds) are in croatia_llocv.npz."""
    from rfsi_b200 import rda
    base = Path("/root/reference/temp_croatia/temp_data")
    g = np.load(OUT / "croatia_llocv.npz")
    xy, temp, fold = g["xy"], g["TEMP"], g["fold"]
    obs = np.asarray(list(rda.read_rda(base / "IDW_10f_obs.rda").values())[0], dtype=np.float64)
    pred = np.asarray(list(rda.read_rda(base / "IDW_10f_pred.rda").values())[0], dtype=np.float64)
    assert len(pred) == 55746 and not np.isnan(pred).any()
    ps = np.round(np.arange(1.0, 3.51, 0.1), 1)
    n_of, p_of, pos = [], [], 0
    for k in range(1, 11):
        val, dev = np.flatnonzero(fold == k), np.flatnonzero(fold != k)
        D, Z, rows = [], [], []
        for d in range(temp.shape[0]):
            tr, te = dev[~np.isnan(temp[d, dev])], val[~np.isnan(temp[d, val])]
            if len(te):
                dist, ob, _, _ = oracle.near_obs(xy[te], xy[tr], temp[d, tr], 40)
                D.append(dist); Z.append(ob); rows.append(temp[d, te])
        o, Dm, Zm = np.concatenate(rows), np.concatenate(D), np.concatenate(Z)
        assert np.array_equal(obs[pos:pos + len(o)], o)            # the stored vectors run fold by fold, day by day
        errs = sorted((float(np.max(np.abs((Dm[:, :n] ** -p * Zm[:, :n]).sum(1) / (Dm[:, :n] ** -p).sum(1) - pred[pos:pos + len(o)]))), n, p)
                      for n in range(2, 41) for p in ps)
        assert errs[0][0] < 3e-14 and errs[1][0] > 1e-3, errs[:2]
        n_of.append(errs[0][1]); p_of.append(errs[0][2])
        pos += len(o)
    assert pos == len(pred)
    np.savez_compressed(OUT / "croatia_idw_cv.npz", pred=pred, n=np.array(n_of), p=np.array(p_of))
    print("IDW (n, p) per fold:", list(zip(n_of, p_of)))


if __name__ == "__main__":
    if Path("/root/reference").exists():
        knn_croatia_real()
        croatia_llocv_inputs()
        croatia_idw_cv()
    knn_lattice()
    knn_realgeom()
    forest_small()
    for p in sorted(OUT.glob("*.npz")):
        print(p.name, p.stat().st_size)
