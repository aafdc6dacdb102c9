"""rfsi() / pred_rfsi() call shapes (README.md:93-143) end to end on a real B200, checked
against the CPU oracle run on the very same fitted forest."""
import numpy as np
import pandas as pd
import pytest

import oracle
import rfsi_b200 as rb
from rfsi_b200 import synth
from rfsi_b200.rfsi import pred_rfsi, rfsi
from tests.util import same_mean

pytestmark = pytest.mark.gpu


def _frames(ndays=3, nsta=120, nnew=900, seed=0):
    rng = np.random.default_rng(seed)
    xy = rng.uniform(0, 1000, size=(nsta, 2))
    rows = []
    for d in range(ndays):
        elev = synth.smooth_field(xy, 5, length_scale=300.0)
        z = 10 + 3 * synth.smooth_field(xy, 9 + d, length_scale=250.0) + 2 * elev + rng.normal(0, .3, nsta)
        for s in range(nsta):
            if rng.uniform() < 0.05:
                continue                                   # station silent that day
            rows.append((s + 1, xy[s, 0], xy[s, 1], d, z[s], elev[s]))
    data = pd.DataFrame(rows, columns=["staid", "x", "y", "day", "temp", "elev"])
    nxy = rng.uniform(0, 1000, size=(nnew, 2))
    new = pd.concat([pd.DataFrame({"id": np.arange(nnew), "x": nxy[:, 0], "y": nxy[:, 1], "day": d,
                                   "elev": synth.smooth_field(nxy, 5, length_scale=300.0)}) for d in range(ndays)],
                    ignore_index=True)
    return data, new


def test_rfsi_then_pred_rfsi_matches_oracle():
    data, new = _frames()
    model = rfsi("temp ~ elev", data, data_staid_x_y_z=["staid", "x", "y", "day"], n_obs=5, num_trees=40,
                 quantreg=True, seed=1)
    assert model.feature_names == ["elev"] + [f"dist{j}" for j in range(1, 6)] + [f"obs{j}" for j in range(1, 6)]
    out = pred_rfsi(model, data, "temp", ["staid", "x", "y", "day"], new, ["id", "x", "y", "day"],
                    type="quantiles", quantiles=(0.1, 0.5, 0.9))
    assert list(out.columns) == ["id", "x", "y", "day", "pred", "quantile..0.1", "quantile..0.5", "quantile..0.9"]
    assert len(out) == len(new)
    fo = model.forest.flat.as_dict()
    for d in range(3):
        st = data[data.day == d]
        nd = new[new.day == d]
        ref = oracle.predict_fused(fo, model.feature_names, loc_xy=nd[["x", "y"]].to_numpy(),
                                   obs_xy=st[["x", "y"]].to_numpy(), obs_z=st["temp"].to_numpy()[None, :], n_obs=5,
                                   covariates={"elev": nd["elev"].to_numpy()}, quantiles=[0.1, 0.5, 0.9])
        got = out[out.day == d]
        same_mean(got["pred"].to_numpy(), ref["mean"][0])
        np.testing.assert_array_equal(got["quantile..0.5"].to_numpy(), ref["quantiles"][1, 0])
    # forest means are convex combinations of training responses
    assert out["pred"].between(data["temp"].min(), data["temp"].max()).all()


def test_spatial_only_and_argument_errors():
    data, new = _frames(ndays=1)
    model = rfsi("temp ~ elev", data.drop(columns="day"), data_staid_x_y_z=["staid", "x", "y", None], n_obs=4,
                 num_trees=10)
    out = pred_rfsi(model, data.drop(columns="day"), "temp", ["staid", "x", "y", None], new.drop(columns="day"),
                    ["id", "x", "y", None])
    assert list(out.columns) == ["id", "x", "y", "pred"] and not out["pred"].isna().any()
    with pytest.raises(ValueError):
        pred_rfsi(model, data, "temp", ["staid", "x", "y", None], new, ["id", "x", "y", None], type="quantiles")
    with pytest.raises(NotImplementedError):
        pred_rfsi(model, data, "temp", ["staid", "x", "y", None], new, ["id", "x", "y", None], zero_tol=1.0)


def test_grown_forest_quantiles_and_simulation_call_shape():
    """simulation.R:186-218 in Python clothes: near.obs on the whole grid, cbind, predict."""
    case = synth.make_case("c1")
    loc = case.locations()
    obs_df = pd.DataFrame({"x": case.obs_xy[:, 0], "y": case.obs_xy[:, 1], "sim1": case.obs_z[0]})
    nearest = rb.near_obs(pd.DataFrame(loc, columns=["x", "y"]), obs_df, zcol="sim1", n_obs=6)
    sim_df = pd.concat([pd.DataFrame(loc, columns=["x", "y"]), nearest], axis=1)
    train = rb.near_obs(obs_df, obs_df, zcol="sim1", n_obs=6)
    from rfsi_b200 import grow
    flat = grow.grow_forest(train.to_numpy(), case.obs_z[0], list(train.columns), num_trees=250, quantreg=True)
    model = rb.RangerForest(flat)
    pred = rb.predict(model, sim_df).predictions                    # extra columns x, y are ignored
    same_mean(pred, oracle.forest_predict(flat.as_dict(), nearest.to_numpy(), threads=8))
    q = rb.predict(model, sim_df, type="quantiles", quantiles=[0.25, 0.5, 0.75]).predictions
    np.testing.assert_array_equal(q, oracle.forest_quantiles(flat.as_dict(), nearest.to_numpy(), [0.25, 0.5, 0.75], threads=8))
