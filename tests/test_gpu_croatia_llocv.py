"""End-to-end anchor on the reference's REAL data: leave-location-out 10-fold CV of RFSI on
the Croatian daily temperatures (157 stations x 366 days, the covariates of 1_models.R:73, the
station folds of temp_data/folds.rda), every near.obs and every predict through the CUDA engine.

The reference's stored result for this exact experiment is RMSE 1.3988 degC / CCC 0.9827
(1_models.R:1245-1249, temp_data/rfsi_10f_{obs,pred}.rda).  It cannot be reproduced bit for bit
(ranger's RNG, per-fold hyper-parameter search, R unavailable); what can be checked is that the
same pipeline on the same data lands at the same accuracy -- an end-to-end parity band next to
the bit-level parity tests against the oracle.  Measured: RMSE 1.396 vs the reference's 1.399
(with scikit-learn's exact CART as the trainer instead of csrc/grow.cpp: 1.367)."""
from pathlib import Path

import numpy as np
import pytest

import rfsi_b200 as rb
from rfsi_b200 import grow

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
STATIC = ["HRdem", "HRdsea", "Lat", "Lon", "HRtwi"]
DAILY = ["INSOL", "ctd", "MODIS_LST"]
N_OBS = 7            # dev_parameters$no (1_models.R:1028-1031)


def test_llocv_on_real_croatian_data_reaches_the_published_accuracy():
    g = np.load(GOLD / "croatia_llocv.npz")
    xy, temp, fold = g["xy"], g["TEMP"], g["fold"]
    stat = {n: g[n].astype(np.float64) for n in STATIC}
    day = {n: g[n].astype(np.float64) for n in DAILY}
    D, S = temp.shape
    # complete.cases over TEMP + covariates (1_models.R:940): only those rows exist at all
    complete = ~np.isnan(temp)
    for n in DAILY:
        complete &= ~np.isnan(day[n])
    assert complete.sum() == int(g["ref_n"]) == 50354
    z = np.where(complete, temp, np.nan)
    names = STATIC + DAILY + [f"dist{j + 1}" for j in range(N_OBS)] + [f"obs{j + 1}" for j in range(N_OBS)]
    obs_all, pred_all = [], []
    for k in range(1, 11):
        dev, val = np.flatnonzero(fold != k), np.flatnonzero(fold == k)
        # training table: per day near.obs(dev stations, dev stations) -- one engine call for all days
        dist, ob = rb.near_obs_days(xy[dev], z[:, dev], N_OBS)
        rows = complete[:, dev]
        cols = [np.broadcast_to(stat[n][dev], (D, len(dev)))[rows] for n in STATIC] + [day[n][:, dev][rows] for n in DAILY]
        cols += [dist[:, :, j][rows] for j in range(N_OBS)] + [ob[:, :, j][rows] for j in range(N_OBS)]
        X = np.stack(cols, axis=1)
        keep = ~np.isnan(X).any(axis=1)
        flat = grow.grow_forest(X[keep], z[:, dev][rows][keep], names, num_trees=250, mtry=4, min_node_size=6, seed=42)
        model = rb.RangerForest(flat)
        # validation: the fold's stations, neighbours taken from the dev stations' observations
        cov = {n: stat[n][val] for n in STATIC}
        cov.update({n: np.nan_to_num(day[n][:, val]) for n in DAILY})     # NA rows are dropped below
        res = rb.predict_fused(model, obs_xy=xy[dev], obs_z=z[:, dev], n_obs=N_OBS, locations=xy[val], covariates=cov)
        vr = complete[:, val]
        obs_all.append(temp[:, val][vr]); pred_all.append(res["mean"][vr])
        model.close()
    o, p = np.concatenate(obs_all), np.concatenate(pred_all)
    assert len(o) == 50354 and not np.isnan(p).any()
    rmse = float(np.sqrt(np.mean((o - p) ** 2)))
    ccc = float(2 * np.cov(o, p, bias=True)[0, 1] / (o.var() + p.var() + (o.mean() - p.mean()) ** 2))
    print(f"LLOCV on real data: RMSE {rmse:.4f} (reference {float(g['ref_rmse']):.4f}), CCC {ccc:.4f} (reference 0.9827)")
    assert abs(float(g["ref_rmse"]) - 1.398816) < 1e-6
    assert 0.93 * 1.398816 < rmse < 1.05 * 1.398816 and ccc > 0.98
