"""Host-side logic above the C ABI (no GPU): ranger import, column matching, feature maps."""
import numpy as np
import pandas as pd
import pytest

import rfsi_b200 as rb
from rfsi_b200 import api, synth
from rfsi_b200.forest import flatten_ranger, ranger_dict_from_flat


def _ranger_like():
    return {
        "num.trees": 2, "treetype": "Regression",
        "forest": {
            "num.trees": 2,
            "child.nodeIDs": [([1, 0, 0], [2, 0, 0]), ([1, 3, 0, 0, 0], [2, 4, 0, 0, 0])],
            "split.varIDs": [[1, 0, 0], [0, 2, 0, 0, 0]],
            "split.values": [[0.5, 1.0, 2.0], [3.0, 7.0, 9.0, 4.0, 5.0]],
            "independent.variable.names": ["a", "b", "c"],
            "is.ordered": [True, True, True],
        },
        "random.node.values": np.array([[np.nan, np.nan], [1.5, np.nan], [2.5, 9.5], [np.nan, 4.5], [np.nan, 5.5]]),
    }


def test_flatten_ranger_roundtrip():
    ff = flatten_ranger(_ranger_like())
    assert ff.ntrees == 2 and ff.node_offsets.tolist() == [0, 3, 8]
    assert ff.left.tolist() == [1, 0, 0, 1, 3, 0, 0, 0] and ff.split_var.tolist() == [1, 0, 0, 0, 2, 0, 0, 0]
    assert ff.rnv_offsets.tolist() == [0, 5]
    assert ff.random_node_values[5 + 2] == 9.5 and np.isnan(ff.random_node_values[0])
    back = ranger_dict_from_flat(ff)
    ff2 = flatten_ranger(back)
    for k in ("node_offsets", "left", "right", "split_var", "split_val"):
        np.testing.assert_array_equal(getattr(ff, k), getattr(ff2, k))


def test_old_ranger_dependent_varid_is_rebased():
    m = _ranger_like()
    # old objects index columns of the training frame including the response at dependent.varID
    m["forest"]["dependent.varID"] = 1
    m["forest"]["split.varIDs"] = [[2, 0, 0], [0, 3, 0, 0, 0]]
    ff = flatten_ranger(m)
    assert ff.split_var.tolist() == [1, 0, 0, 0, 2, 0, 0, 0]


def test_unsupported_models_are_refused():
    m = _ranger_like(); m["treetype"] = "Classification"
    with pytest.raises(ValueError):
        flatten_ranger(m)
    m = _ranger_like(); m["forest"]["is.ordered"] = [True, False, True]
    with pytest.raises(ValueError):
        flatten_ranger(m)


def test_model_matrix_matches_by_name_and_ignores_extras():
    model = rb.RangerForest(flatten_ranger(_ranger_like()))
    df = pd.DataFrame({"c": [1.0, 2.0], "zzz": [9.0, 9.0], "a": [3.0, 4.0], "b": [5.0, 6.0]})
    Xt = api._model_matrix(model, df)
    assert Xt.shape == (3, 2) and Xt.tolist() == [[3.0, 4.0], [5.0, 6.0], [1.0, 2.0]]
    with pytest.raises(ValueError):
        api._model_matrix(model, df.drop(columns=["b"]))
    with pytest.raises(ValueError):
        api._model_matrix(model, np.zeros((4, 2)))


def test_feature_map_codes():
    names = ["cov1", "dist1", "dist2", "obs1", "obs2", "elev"]
    fm = api.feature_map(names, ["elev", "cov1"], 2)
    assert fm.tolist() == [1, 0x10000, 0x10001, 0x20000, 0x20001, 0]
    with pytest.raises(ValueError):
        api.feature_map(["dist3"], [], 2)


def test_near_obs_argument_errors_mirror_reference_usage():
    with pytest.raises(NotImplementedError):
        rb.near_obs(np.zeros((2, 2)), np.zeros((2, 3)), idw=True)
    with pytest.raises(ValueError):
        rb.near_obs(np.zeros((2, 2)), np.zeros((2, 3)), n_obs=0)


def test_synth_cases_have_the_named_shapes():
    c1 = synth.make_case("c1")
    assert c1.npix == 10000 and c1.obs_xy.shape == (200, 2) and c1.n_obs == 6 and len(c1.feature_names) == 12
    # stations are pixel centres: self matches exist
    loc = c1.locations()
    assert set(map(tuple, c1.obs_xy)).issubset(set(map(tuple, loc)))
    c4 = synth.make_case("c4")
    assert c4.npix == 380 * 280 and c4.obs_xy.shape[0] == 87 and len(c4.feature_names) == 17
    miss = np.isnan(c4.obs_z).sum(axis=1)
    assert miss.min() >= 1 and miss.max() <= 5
    # raster form and explicit form of the coordinates agree bit for bit
    c3 = synth.make_case("c3", ndays=1)
    x0, y0, dx, dy, ncol = c3.grid
    p = np.arange(1000, 1500)
    np.testing.assert_array_equal(c3.locations(1000, 500)[:, 0], x0 + (p % ncol) * dx)


def _stump_model():
    from rfsi_b200.forest import FlatForest
    return rb.RangerForest(FlatForest(1, np.array([0, 3], dtype=np.int64), np.array([1, 0, 0], dtype=np.int32),
                                      np.array([2, 0, 0], dtype=np.int32), np.array([0, 0, 0], dtype=np.int32),
                                      np.array([0.5, 1.0, 2.0]), ["tmin"]))


def test_predict_fused_checks_its_python_level_arguments_before_the_engine():
    """Names and shapes are resolved on the host: a wrong one is a ValueError here, not an engine status."""
    model = _stump_model()
    loc = np.zeros((4, 2))
    kw = dict(obs_xy=np.zeros((3, 2)), obs_z=np.zeros((2, 3)), n_obs=1, locations=loc)
    with pytest.raises(ValueError, match="neither dist/obs"):
        rb.predict_fused(model, covariates={"tmax": np.zeros(4)}, **kw)                    # model wants tmin
    with pytest.raises(ValueError, match="shape"):
        rb.predict_fused(model, covariates={"tmin": np.zeros(5)}, **kw)
    with pytest.raises(ValueError, match="cov_fix"):
        rb.predict_fused(model, covariates={"tmin": np.zeros(4)}, cov_fix=[("tmin", "tmax", 1.0)], **kw)
    with pytest.raises(ValueError, match="cov_locations"):
        rb.predict_fused(model, covariates={"tmin": np.zeros(4)}, cov_locations=np.zeros((3, 2)), **kw)
    with pytest.raises(ValueError, match="nobs"):
        rb.predict_fused(model, covariates={"tmin": np.zeros(4)}, obs_xy=np.zeros((2, 2)), obs_z=np.zeros((2, 3)),
                         n_obs=1, locations=loc)
    with pytest.raises(ValueError, match="raster dtype"):
        rb.predict_fused(model, covariates={"tmin": rb.Raster(np.zeros((2, 2), dtype=np.int64), 0, 2, 1, 1)}, **kw)


def test_engine_validates_repairs_and_device_lists_on_the_host(has_gpu):
    """rfsi_predict_fused(_multi): argument errors are found before any device work (RFSI_E_ARG),
    so they show even on a box without a GPU, where everything else is RFSI_E_NO_DEVICE."""
    import ctypes as C
    from rfsi_b200 import _lib
    lib = _lib.load()
    model = _stump_model()
    loc = np.zeros((4, 2))
    a = _lib.FusedArgs()
    a.struct_size = C.sizeof(_lib.FusedArgs)
    x = np.zeros(4); z = np.zeros((1, 3)); fmap = np.array([0], dtype=np.int32)
    ptr = lambda v: v.ctypes.data_as(C.c_void_p)
    a.loc_x, a.loc_y, a.npix, a.grid_ncol = ptr(x), ptr(x), 4, 1
    a.obs_x, a.obs_y, a.obs_z, a.nobs, a.ndays, a.n_obs = ptr(x), ptr(x), ptr(z), 3, 1, 1
    cols = (C.c_void_p * 1)(x.ctypes.data); strides = (C.c_int64 * 1)(0)
    a.ncov, a.cov, a.cov_day_stride = 1, C.cast(cols, C.POINTER(C.c_void_p)), C.cast(strides, C.POINTER(C.c_int64))
    a.feat_map = fmap.ctypes.data_as(C.POINTER(C.c_int32))
    out = np.zeros(4); a.out_mean = ptr(out)
    fix = (_lib.CovFix * 1)(_lib.CovFix(0, 0, 1.0))              # a covariate repaired against itself
    a.cov_fix, a.ncov_fix = C.cast(fix, C.POINTER(_lib.CovFix)), 1
    assert lib.rfsi_predict_fused(model.handle, C.byref(a), 0) == _lib.RFSI_E_ARG
    assert b"cov_fix" in lib.rfsi_last_error()
    a.ncov_fix = 0
    a.cov_loc_x = ptr(x)                                         # x without y
    assert lib.rfsi_predict_fused(model.handle, C.byref(a), 0) == _lib.RFSI_E_ARG
    a.cov_loc_x = None
    devs = (C.c_int * 2)(0, 0)
    assert lib.rfsi_predict_fused_multi(model.handle, C.byref(a), devs, 0) == _lib.RFSI_E_ARG
    assert lib.rfsi_predict_fused_multi(model.handle, C.byref(a), None, 2) == _lib.RFSI_E_ARG
    rc = lib.rfsi_predict_fused_multi(model.handle, C.byref(a), devs, 2)
    assert rc == (0 if has_gpu else _lib.RFSI_E_NO_DEVICE)
    a.struct_size = 8                                            # header / library mismatch
    assert lib.rfsi_predict_fused_multi(model.handle, C.byref(a), devs, 2) == _lib.RFSI_E_ARG


def test_pred_rfsi_with_a_raster_newdata_wires_grid_mask_and_layers(monkeypatch):
    """README.md:121-153: newdata <- terra::rast(...), output.format = "SpatRaster".  The engine call is
    replaced by a recorder here (the GPU side is the fused call of tests/test_gpu_raster.py): what is
    checked is the host wiring -- cell centres in row-major order, the NA mask over all covariate
    layers, layers handed over as rasters, results put back on the grid."""
    from rfsi_b200 import rfsi as R
    seen = {}

    def fake(model, **kw):
        seen.update(kw)
        n = len(kw["locations"])
        out = {"mean": np.arange(n, dtype=np.float64)[None, :]}
        if kw["quantiles"] is not None:
            out["quantiles"] = np.stack([np.full((1, n), q) for q in kw["quantiles"]])
        return out
    monkeypatch.setattr(R.api, "predict_fused", fake)
    dist = np.array([[0.1, 0.2, np.nan], [0.4, 0.5, 0.6]])
    soil = np.array([[1, 2, 3], [-9, 2, 1]], dtype=np.int16)
    newdata = {"dist": rb.Raster(dist, 178000.0, 334000.0, 40.0, 40.0), "soil": rb.Raster(soil, 178000.0, 334000.0, 40.0, 40.0, nodata=-9),
               "unused": rb.Raster(np.full((2, 3), np.nan), 178000.0, 334000.0, 40.0, 40.0)}
    model = R.RfsiModel(forest=None, formula="zinc ~ dist + soil", response="zinc", covariates=["dist", "soil"], n_obs=5,
                        num_trees=250, quantreg=True)
    data = pd.DataFrame({"id": [1, 2, 3], "x": [178010.0, 178050.0, 178090.0], "y": [333990.0, 333950.0, 333970.0],
                         "zinc": [100.0, np.nan, 300.0]})
    out = R.pred_rfsi(model, data, "zinc", ("id", "x", "y", None), newdata, None, output_format="SpatRaster",
                      type="quantiles", quantiles=(0.1, 0.9))
    np.testing.assert_array_equal(seen["locations"], [[178020, 333980], [178060, 333980], [178100, 333980],
                                                      [178020, 333940], [178060, 333940], [178100, 333940]])
    np.testing.assert_array_equal(seen["pix_mask"], [1, 1, 0, 0, 1, 1])           # NA in dist, NAvalue in soil; `unused` is not a model column
    assert list(seen["covariates"]) == ["dist", "soil"] and seen["n_obs"] == 5
    np.testing.assert_array_equal(seen["obs_z"], [[100.0, np.nan, 300.0]])
    assert out.names() == ["pred", "quantile..0.1", "quantile..0.9"]
    np.testing.assert_array_equal(out["pred"], [[0, 1, 2], [3, 4, 5]])
    assert out["quantile..0.9"].shape == (2, 3) and (out.x_ul, out.y_ul, out.dx, out.dy) == (178000.0, 334000.0, 40.0, 40.0)
    with pytest.raises(NotImplementedError):
        R.pred_rfsi(model, data, "zinc", ("id", "x", "y", None), newdata, None, output_format="data.frame")
    bad = dict(newdata, soil=rb.Raster(soil[:, :2], 178000.0, 334000.0, 40.0, 40.0))
    with pytest.raises(ValueError, match="grid"):
        R.pred_rfsi(model, data, "zinc", ("id", "x", "y", None), bad, None, output_format="SpatRaster")
    with pytest.raises(ValueError, match="no layer"):
        R.pred_rfsi(model, data, "zinc", ("id", "x", "y", None), {"dist": newdata["dist"]}, None, output_format="SpatRaster")
