"""The R-free .rda/.rds reader: round trips of ranger-shaped model lists through real
RDX2/RDX3 files, and (in the build container only, where /root/reference is mounted) the
reference's own .rda files, whose contents are known from the scripts' comments."""
from pathlib import Path

import numpy as np
import pytest

from rfsi_b200 import rda
from rfsi_b200.forest import flatten_ranger

REF = Path("/root/reference")


def _ranger_like(T=3, seed=0):
    rng = np.random.default_rng(seed)
    child, var, val = [], [], []
    for t in range(T):
        n = 2 * int(rng.integers(2, 6)) + 1
        left = np.zeros(n, dtype=np.int32); right = left.copy()
        nxt = 1
        for i in range(n):
            if nxt + 1 < n:
                left[i], right[i] = nxt, nxt + 1
                nxt += 2
        child.append([left, right])
        var.append(rng.integers(0, 4, n).astype(np.int32))
        val.append(rng.normal(size=n))
    m = max(len(v) for v in var)
    rnv = np.full((m, T), np.nan)
    for t in range(T):
        rnv[: len(var[t]), t] = rng.normal(size=len(var[t]))
    return {
        "predictions": rng.normal(size=5),
        "num.trees": np.array([T], dtype=np.int32),
        "forest": {"num.trees": np.array([T], dtype=np.int32), "child.nodeIDs": child, "split.varIDs": var,
                   "split.values": val, "is.ordered": np.array([True] * 4),
                   "independent.variable.names": ["dist1", "obs1", "elev", "twi"]},
        "random.node.values": rnv,
        "treetype": "Regression",
    }


@pytest.mark.parametrize("version,compress", [(3, "gzip"), (2, "gzip"), (3, "bzip2"), (3, "xz"), (3, "none")])
def test_ranger_model_round_trip(tmp_path, version, compress):
    model = _ranger_like()
    path = tmp_path / "RFSI.rda"
    rda.write_rda(path, {"rfsi_model": model}, version=version, compress=compress, classes={"rfsi_model": "ranger"})
    ws = rda.read_rda(path)
    assert list(ws) == ["rfsi_model"] and list(ws["rfsi_model"].attrs["class"]) == ["ranger"]
    got = rda.load_ranger(path)
    assert got["treetype"] == "Regression" and got["forest"]["independent.variable.names"] == ["dist1", "obs1", "elev", "twi"]
    a = flatten_ranger(got)
    b = flatten_ranger({**model, "forest": {**model["forest"], "num.trees": 3,
                                             "child.nodeIDs": [tuple(c) for c in model["forest"]["child.nodeIDs"]]}})
    for k in ("node_offsets", "left", "right", "split_var", "split_val", "rnv_offsets"):
        np.testing.assert_array_equal(getattr(a, k), getattr(b, k))
    np.testing.assert_array_equal(a.random_node_values, b.random_node_values)      # NaN positions included


def test_two_rangers_need_a_name(tmp_path):
    path = tmp_path / "two.rda"
    rda.write_rda(path, {"a": _ranger_like(seed=1), "b": _ranger_like(seed=2), "n_obs": np.array([10], dtype=np.int32)},
                  classes={"a": "ranger", "b": "ranger"})
    with pytest.raises(ValueError):
        rda.load_ranger(path)
    assert rda.load_ranger(path, "b")["forest"]["num.trees"] == 3
    assert int(rda.read_rda(path)["n_obs"][0]) == 10


@pytest.mark.skipif(not REF.exists(), reason="the reference tree is only mounted in the build container")
def test_reads_the_references_own_rda_files():
    # simulation/accuracy_results.rda: 133 200 x 7 character matrix written by simulation.R:197-215;
    # mean RFSI "DT" at 200 points is 1.482 s
    ws = rda.read_rda(REF / "simulation" / "accuracy_results.rda")
    (name, m), = ws.items()
    m = np.array(m, dtype=object).reshape(tuple(int(d) for d in m.attrs["dim"]), order="F")
    assert m.shape == (133200, 7)
    sel = (m[:, 1] == "DT") & (m[:, 6] == "RFSI") & (m[:, 4] == "200")
    assert sel.sum() == 600
    assert abs(np.mean([float(v) for v in m[sel, 0]]) - 1.482) < 5e-3
    # temp_croatia station folds: 10 folds over 157 stations
    strat = rda.read_rda(REF / "temp_croatia" / "temp_data" / "folds.rda")["strat"]
    assert strat.names == ["data", "obs.fold.list"]
    assert [len(f) for f in strat["obs.fold.list"]] == [20, 17, 11, 17, 19, 14, 12, 17, 12, 18]
    assert np.asarray(strat["data"]["fold"]).shape == (157,)
    # an S4 object (STFDF: 157 stations x 366 days) parses too
    st = rda.read_rda(REF / "temp_croatia" / "temp_data" / "stfdf.rda")["stfdf"]
    assert st.kind == "S4" and list(st.attrs["class"]) == ["STFDF"]
    assert st.attrs["sp"].attrs["coords"].shape == (157, 2)
    temp = np.asarray(st.attrs["data"]["TEMP"])
    assert temp.shape == (157 * 366,) and int(np.isnan(temp).sum()) == 1716
