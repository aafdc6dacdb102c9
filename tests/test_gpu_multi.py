"""rfsi_predict_fused_multi: one call, the raster cut into one band of pixels per device, one host
thread per band (north_star (c): no collective on the compute path, every band's D2H copies
land in its slice of the caller's arrays).  The banding must not change a single bit of the
result.  On a one-GPU box the bands share device 0 (devices may repeat), which exercises the
banding, the threading and the slice arithmetic; with two or more GPUs the same checks run
across real devices."""
import warnings

import numpy as np
import pytest

import oracle
import rfsi_b200 as rb
from rfsi_b200 import grow, synth
from tests.util import same_mean

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c4():
    case = synth.make_case("c4", ndays=3)
    X, y = synth.training_table(case, lambda l, o, z, n: oracle.near_obs(l, o, z, n)[:2])
    flat = grow.grow_forest(X, y, case.feature_names, num_trees=40, quantreg=True)
    loc = case.locations()
    cov = {n: np.stack([f(loc, d) for d in range(3)]) for n, f in case.cov_daily.items()}
    return case, flat, rb.RangerForest(flat), loc, cov


def _device_sets():
    n = rb.device_count()
    sets = [[0, 0], [0, 0, 0, 0, 0]]
    if n >= 2:
        sets += [list(range(n)), [1, 0]]
    return sets


def _same(a, b):
    assert a.keys() == b.keys()
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


def test_bands_reproduce_the_single_device_call(c4):
    case, flat, model, loc, cov = c4
    kw = dict(obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, covariates=cov, quantiles=[0.25, 0.5, 0.75],
              centi_int16=True)
    mask = (np.arange(case.npix) % 3 != 0).astype(np.uint8)
    perm = np.random.default_rng(0).permutation(case.npix)
    forms = {
        "grid": dict(grid=case.grid, npix=case.npix),
        "grid, offset band + mask": dict(grid=case.grid, npix=case.npix - 5 * case.ncol - 17, pix_begin=5 * case.ncol, pix_mask=mask[:case.npix - 5 * case.ncol - 17]),
        "explicit rows": dict(locations=loc),
        "explicit shuffled": dict(locations=loc[perm]),
    }
    for name, form in forms.items():
        k = dict(kw)
        n = form.get("npix", case.npix)
        b = form.get("pix_begin", 0)
        if name == "explicit shuffled":
            k["covariates"] = {c: v[:, perm] for c, v in cov.items()}
        elif n != case.npix or b:
            k["covariates"] = {c: v[:, b:b + n] for c, v in cov.items()}
        one = rb.predict_fused(model, **k, **form)
        for devs in _device_sets():
            _same(one, rb.predict_fused(model, devices=devs, **k, **form))
    ref = oracle.predict_fused(flat.as_dict(), case.feature_names, loc_xy=loc[:1500], obs_xy=case.obs_xy, obs_z=case.obs_z,
                               n_obs=case.n_obs, covariates={c: v[:, :1500] for c, v in cov.items()})
    got = rb.predict_fused(model, devices=[0, 0, 0], locations=loc[:1500], covariates={c: v[:, :1500] for c, v in cov.items()},
                           obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs)
    same_mean(got["mean"], ref["mean"])


def test_more_bands_than_rows_and_tiny_inputs(c4):
    case, flat, model, loc, cov = c4
    kw = dict(obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs)
    for n in (1, 2, 7, case.ncol, case.ncol + 1):
        c = {k: v[:, :n] for k, v in cov.items()}
        one = rb.predict_fused(model, locations=loc[:n], covariates=c, **kw)
        _same(one, rb.predict_fused(model, locations=loc[:n], covariates=c, devices=[0] * 6, **kw))
        one = rb.predict_fused(model, grid=case.grid, npix=n, covariates=c, **kw)
        _same(one, rb.predict_fused(model, grid=case.grid, npix=n, covariates=c, devices=[0] * 6, **kw))


def test_warnings_and_errors_of_any_band_reach_the_caller(c4):
    case, flat, model, loc, cov = c4
    z = case.obs_z.copy()
    z[1, 3:] = np.nan                                   # day 2: three stations for n.obs = 7 -> NA outputs + warning
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        one = rb.predict_fused(model, obs_xy=case.obs_xy, obs_z=z, n_obs=case.n_obs, locations=loc, covariates=cov)
        many = rb.predict_fused(model, obs_xy=case.obs_xy, obs_z=z, n_obs=case.n_obs, locations=loc, covariates=cov,
                                devices=[0, 0, 0])
    assert len(w) == 2 and "fewer valid stations" in str(w[1].message)
    _same(one, many)
    assert np.isnan(many["mean"][1]).all() and not np.isnan(many["mean"][0]).any()
    bad = {k: v.copy() for k, v in cov.items()}
    bad["tmax"][2, case.npix - 5] = np.nan              # NA in a model column, last band only
    with pytest.raises(rb.RfsiError) as e:
        rb.predict_fused(model, obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, locations=loc, covariates=bad,
                         devices=[0, 0, 0])
    assert e.value.code == -4
    with pytest.raises(rb.RfsiError) as e:
        rb.predict_fused(model, obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, locations=loc, covariates=cov,
                         devices=[0, 99])
    assert e.value.code == -2


def test_rasters_and_repairs_across_bands(c4):
    case, flat, model, loc, cov = c4
    rng = np.random.default_rng(2)
    nrow, ncol = case.nrow, case.ncol
    tmax = rng.integers(-50, 350, size=(3, nrow, ncol)).astype(np.int16)
    tmin = (tmax - rng.integers(-30, 120, size=tmax.shape)).astype(np.int16)
    geo = dict(x_ul=0.0, y_ul=float(nrow), dx=1.0, dy=1.0)
    R = {"tmax": rb.Raster(tmax, nodata=-32767, scale=10.0, divide=True, **geo),
         "tmin": rb.Raster(tmin, nodata=-32767, scale=10.0, divide=True, **geo), "imerg": cov["imerg"]}
    kw = dict(obs_xy=case.obs_xy, obs_z=case.obs_z, n_obs=case.n_obs, covariates=R, cov_fix=[("tmin", "tmax", 1.0)],
              locations=loc, cov_locations=loc.copy(), quantiles=[0.1, 0.9])
    _same(rb.predict_fused(model, **kw), rb.predict_fused(model, devices=[0, 0, 0], **kw))
