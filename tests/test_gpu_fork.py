"""Forked workers, the way the reference gets its parallelism (registerDoParallel + foreach %dopar%:
simulation/simulation.R:113-114 over seeds, prcp_catalonia/5_prcp_prediction.R:243-244 over days): the model is
imported ONCE in the parent (rfsi_forest_create touches no device), every forked child makes its own first device
call and maps its own days; results are bit-equal to the oracle.  And the case that cannot work -- the parent used
the GPU before forking -- is refused with RFSI_E_FORKED and a message that says what to do, not with the CUDA
runtime's "initialization error".  Runs in a fresh interpreter: the pytest process itself has used CUDA already."""
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

SCRIPT = textwrap.dedent('''
    import os, sys, pickle
    import numpy as np
    sys.path.insert(0, %r)
    import oracle
    import rfsi_b200 as rb
    from rfsi_b200 import synth, _lib

    touch_first = sys.argv[1] == "touch"
    case = synth.make_case("c1", ndays=4)
    case.obs_z[2, ::5] = np.nan
    X, y = synth.training_table(case, lambda l, o, z, n: oracle.near_obs(l, o, z, n)[:2])
    flat = synth.grow_forest_sklearn(X, y, case.feature_names, num_trees=30, quantreg=True, n_jobs=1)
    model = rb.RangerForest(flat)            # host-side import only: no device call yet
    if touch_first:
        assert rb.device_count() >= 1        # the parent initialises CUDA ...
    kids = []
    for day in range(4):
        r, w = os.pipe()
        pid = os.fork()                      # ... and forks its day workers (5_prcp_prediction.R:243-244)
        if pid == 0:
            os.close(r)
            try:
                got = rb.predict_fused(model, obs_xy=case.obs_xy, obs_z=case.obs_z[day:day + 1], n_obs=case.n_obs,
                                       grid=case.grid, npix=case.npix, quantiles=[0.25, 0.5, 0.75])
                out = ("ok", got["mean"], got["quantiles"])
            except rb.RfsiError as e:
                out = ("err", e.code, str(e))
            os.write(w, pickle.dumps(out)); os.close(w)
            os._exit(0)                      # no interpreter teardown in the child
        os.close(w)
        kids.append((pid, r, day))
    results = {}
    for pid, r, day in kids:
        buf = b""
        while True:
            chunk = os.read(r, 1 << 20)
            if not chunk:
                break
            buf += chunk
        os.waitpid(pid, 0)
        results[day] = pickle.loads(buf)
    if touch_first:
        for day, res in results.items():
            assert res[0] == "err" and res[1] == _lib.RFSI_E_FORKED, res
            assert "forked" in res[2] and "fork()" in res[2], res[2]
        print("REFUSED", results[0][2])
    else:
        ref = oracle.predict_fused(flat.as_dict(), case.feature_names, loc_xy=case.locations(), obs_xy=case.obs_xy,
                                   obs_z=case.obs_z, n_obs=case.n_obs, quantiles=[0.25, 0.5, 0.75])
        for day, res in results.items():
            assert res[0] == "ok", res
            assert np.array_equal(res[1][0], ref["mean"][day]), day
            assert np.array_equal(res[2][:, 0], ref["quantiles"][:, day], equal_nan=True), day
        # the parent can still make its own first device call afterwards
        got = rb.predict_fused(model, obs_xy=case.obs_xy, obs_z=case.obs_z[:1], n_obs=case.n_obs, grid=case.grid, npix=case.npix)
        assert np.array_equal(got["mean"][0], ref["mean"][0])
        print("BITEQUAL", len(results))
''') % str(ROOT)


def _run(mode):
    return subprocess.run([sys.executable, "-c", SCRIPT, mode], capture_output=True, text=True, timeout=600)


def test_forked_children_each_get_their_own_context():
    r = _run("clean")
    assert r.returncode == 0 and "BITEQUAL 4" in r.stdout, r.stdout + r.stderr


def test_fork_after_the_parent_used_the_gpu_is_refused_clearly():
    r = _run("touch")
    assert r.returncode == 0 and "REFUSED" in r.stdout, r.stdout + r.stderr
