"""Stage-2 pin, ready for the one input this image cannot produce: a ranger forest saved by R
next to R's own predict() output (tests/golden/make_ranger_fixture.R writes
tests/golden/ranger_fixture.rda on a machine with R + ranger + meteo).  When the file is there,
the oracle (here) and the CUDA engine (-m gpu) are held to R's numbers: near.obs distances /
observations, forest means, type-7 quantiles and terminal nodes.  Until then the checks are
exercised on a stand-in file of the same layout written by rfsi_b200/rda.py from the oracle's own
outputs, so the plumbing (reading the workspace, matching columns by name, matrix layouts) is
known to work the day the real file arrives."""
from pathlib import Path

import numpy as np
import pytest

import oracle
from rfsi_b200 import grow, rda
from rfsi_b200.forest import flatten_ranger, ranger_dict_from_flat

FIXTURE = Path(__file__).parent / "golden" / "ranger_fixture.rda"
RTOL = 1e-12          # north_star: predictions within 1e-12 relative in fp64


def load_fixture(path):
    ws = rda.read_rda(path)
    model = rda.ranger_to_dict(ws["model"])
    flat = flatten_ranger(model)

    def frame(df):
        return {n: np.asarray(df[n], dtype=np.float64) for n in df.names}

    def matrix(m):
        return np.asarray(m, dtype=np.float64).reshape([int(d) for d in m.attrs["dim"]], order="F")

    G, gno, st = frame(ws["G"]), frame(ws["gno"]), frame(ws["stations"])
    X = np.stack([G[n] for n in flat.feature_names], axis=1)          # predict() matches columns by name
    return dict(flat=flat, X=X, grid=np.column_stack([G["x"], G["y"]]), gno=gno,
                stations=np.column_stack([st[k] for k in list(st)[:3]]),
                pred=np.asarray(ws["pred"], dtype=np.float64), q=matrix(ws["q"]), tn=matrix(ws["tn"]).astype(np.int64))


def check(fx, near_obs, predict, quantiles, terminal_nodes):
    n = sum(1 for k in fx["gno"] if k.startswith("dist"))
    dist, obs = near_obs(fx["grid"], fx["stations"][:, :2], fx["stations"][:, 2], n)
    for j in range(n):
        np.testing.assert_allclose(dist[:, j], fx["gno"][f"dist{j + 1}"], rtol=RTOL, atol=0)
        np.testing.assert_array_equal(obs[:, j], fx["gno"][f"obs{j + 1}"])
    np.testing.assert_allclose(predict(fx["flat"], fx["X"]), fx["pred"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(quantiles(fx["flat"], fx["X"], [0.25, 0.5, 0.75]), fx["q"], rtol=RTOL, atol=0)
    np.testing.assert_array_equal(terminal_nodes(fx["flat"], fx["X"]), fx["tn"])


def _oracle_fns():
    return (lambda l, o, z, n: oracle.near_obs(l, o, z, n)[:2],
            lambda f, X: oracle.forest_predict(f.as_dict(), X),
            lambda f, X, p: oracle.forest_quantiles(f.as_dict(), X, p),
            lambda f, X: oracle.forest_terminal_nodes(f.as_dict(), X))


def _stand_in(path):
    """A workspace with the objects and layouts make_ranger_fixture.R saves, filled by the oracle."""
    rng = np.random.default_rng(3)
    st = np.column_stack([rng.uniform(4e5, 8e5, 120), rng.uniform(4.7e6, 5.1e6, 120), rng.normal(12, 5, 120)])
    n = 7
    names = ["x", "y"] + [f"dist{j + 1}" for j in range(n)] + [f"obs{j + 1}" for j in range(n)]
    d, o, _, _ = oracle.near_obs(st[:, :2], st[:, :2], st[:, 2], n)
    flat = grow.grow_forest(np.column_stack([st[:, :2], d, o]), st[:, 2], names, num_trees=30, mtry=4, min_node_size=6,
                            seed=42, quantreg=True)
    gx, gy = np.meshgrid(np.linspace(4e5, 8e5, 25), np.linspace(4.7e6, 5.1e6, 25))
    grid = np.column_stack([gx.ravel(), gy.ravel()])
    gd, go, _, _ = oracle.near_obs(grid, st[:, :2], st[:, 2], n)
    gno = {**{f"dist{j + 1}": gd[:, j] for j in range(n)}, **{f"obs{j + 1}": go[:, j] for j in range(n)}}
    G = {"x": grid[:, 0], "y": grid[:, 1], **gno}
    X = np.column_stack([G[k] for k in names])
    fo = flat.as_dict()
    rda.write_rda(path, {"model": ranger_dict_from_flat(flat), "G": G, "gno": gno,
                         "pred": oracle.forest_predict(fo, X), "q": oracle.forest_quantiles(fo, X, [0.25, 0.5, 0.75]),
                         "tn": oracle.forest_terminal_nodes(fo, X).astype(np.int32),
                         "stations": {"x": st[:, 0], "y": st[:, 1], "TEMP": st[:, 2]}},
                  version=2, classes={"model": "ranger", "G": "data.frame", "gno": "data.frame", "stations": "data.frame"})


def test_fixture_plumbing_on_a_stand_in_file(tmp_path):
    _stand_in(tmp_path / "ranger_fixture.rda")
    fx = load_fixture(tmp_path / "ranger_fixture.rda")
    assert fx["X"].shape == (625, 16) and fx["q"].shape == (625, 3) and fx["tn"].shape == (625, 30)
    check(fx, *_oracle_fns())
    fx["pred"] = fx["pred"] + 1e-9                      # the comparison is live
    with pytest.raises(AssertionError):
        check(fx, *_oracle_fns())


@pytest.mark.skipif(not FIXTURE.exists(), reason="tests/golden/ranger_fixture.rda not minted yet (needs R: make_ranger_fixture.R)")
def test_oracle_against_r_ranger_fixture():
    check(load_fixture(FIXTURE), *_oracle_fns())


@pytest.mark.gpu
def test_cuda_engine_against_the_fixture_layout(tmp_path):
    """The same checks through the C ABI: on R's file when it exists, else on the stand-in."""
    import rfsi_b200 as rb
    path = FIXTURE
    if not FIXTURE.exists():
        path = tmp_path / "ranger_fixture.rda"
        _stand_in(path)
    fx = load_fixture(path)
    model = rb.RangerForest(fx["flat"])

    def near(l, o, z, n):
        df = rb.near_obs(l, np.column_stack([o, z]), zcol=2, n_obs=n).to_numpy()
        return df[:, :n], df[:, n:]
    check(fx, near, lambda f, X: rb.predict(model, X).predictions,
          lambda f, X, p: rb.predict(model, X, type="quantiles", quantiles=p).predictions,
          lambda f, X: rb.predict(model, X, type="terminalNodes").predictions)
    model.close()
