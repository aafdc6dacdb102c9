"""R is not installed here.  The .Call shim (rfsi_b200/R/src/r_shim.c) is (1) syntax-checked against
a stub of R's C API, (2) checked for a complete call table, and (3) linked against a working miniature
of that API (tests/tools/mock_libR.c) and RUN: here the entries that need no GPU (forest import, its
finalizer, argument checks, the no-device refusal surfacing as an R error); tests/test_gpu_r_shim.py
runs the compute entries on the GPU against the oracle."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests.rmock import MockR, RError

ROOT = Path(__file__).resolve().parents[1]
RDIR = ROOT / "rfsi_b200" / "R"


def test_r_shim_compiles_against_stub_headers():
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-I", str(ROOT / "include"), "-I", str(RDIR / "src" / "stub"),
                        str(RDIR / "src" / "r_shim.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_call_table_covers_r_level_calls():
    shim = (RDIR / "src" / "r_shim.c").read_text()
    registered = set(re.findall(r'\{"(C_rfsi_[a-z_]+)"', shim))
    used = set()
    for f in (RDIR / "R").glob("*.R"):
        used |= set(re.findall(r"\.Call\((C_rfsi_[a-z_]+)", f.read_text()))
    assert used and used <= registered
    # every engine symbol the shim calls is declared in the public header
    header = (ROOT / "include" / "rfsi_b200.h").read_text()
    for sym in set(re.findall(r"\b(rfsi_[a-z0-9_]+)\s*\(", shim)) - {"rfsi_import_ranger"}:
        assert re.search(rf"\b{sym}\s*\(", header), sym
    # the registered arities are the arities of the C functions
    for name, n in re.findall(r'\{"(C_rfsi_[a-z_]+)", \(DL_FUNC\)&\1, (\d+)\}', shim):
        sig = re.search(rf"SEXP {name}\(([^)]*)\)", shim).group(1)
        assert sig.count("SEXP") == int(n), name


@pytest.fixture(scope="module")
def R():
    r = MockR()
    yield r
    r.gc()


def _stump():
    # two trees: a stump on feature 1 and a single leaf; random.node.values matrix 3 x 2
    off = np.array([0.0, 3.0, 4.0])
    left = np.array([1, 0, 0, 0], dtype=np.int32); right = np.array([2, 0, 0, 0], dtype=np.int32)
    var = np.array([1, 0, 0, 0], dtype=np.int32); val = np.array([0.5, -1.0, 1.0, 10.0])
    rnv = np.array([[np.nan, 3.0], [5.0, np.nan], [7.0, np.nan]])
    return off, left, right, var, val, rnv


def test_forest_import_and_finalizer_run_without_a_gpu(R):
    off, left, right, var, val, rnv = _stump()
    h = R.call("C_rfsi_forest_new", off, left, right, var, val, rnv, 2)
    assert h is not None
    R.lib.mock_finalize(h.sexp)                       # gc() of the handle: the forest is destroyed, the pointer cleared
    with pytest.raises(RError, match="forest handle is NULL"):
        R.call("C_rfsi_predict", h, np.zeros((3, 2)), None, 0)
    with pytest.raises(RError, match="num.trees columns"):
        R.call("C_rfsi_forest_new", off, left, right, var, val, rnv[:, :1], 2)
    bad = left.copy(); bad[0] = 7                     # child id out of range: the engine's message reaches R
    with pytest.raises(RError, match="child id out of range"):
        R.call("C_rfsi_forest_new", off, bad, right, var, val, None, 2)


def test_argument_checks_raise_r_errors(R):
    loc = np.zeros((4, 2)); obs = np.ones((3, 2)); z = np.arange(3.0)
    with pytest.raises(RError, match="bad shapes"):
        R.call("C_rfsi_near_obs", loc, obs, z[:2], 2, True, 0)
    with pytest.raises(RError, match="numeric inputs expected"):
        R.call("C_rfsi_near_obs", loc.astype(np.int32), obs, z, 2, True, 0)
    off, left, right, var, val, rnv = _stump()
    h = R.call("C_rfsi_forest_new", off, left, right, var, val, None, 2)
    with pytest.raises(RError, match="npix x nfeat"):
        R.call("C_rfsi_predict", h, np.zeros((3, 5)), None, 0)


def test_no_device_is_an_r_error_not_a_cpu_fallback(R, has_gpu):
    if has_gpu:
        pytest.skip("a device is visible")
    off, left, right, var, val, rnv = _stump()
    h = R.call("C_rfsi_forest_new", off, left, right, var, val, rnv, 2)
    with pytest.raises(RError, match="no CUDA device|no CPU path"):
        R.call("C_rfsi_predict", h, np.zeros((3, 2)), None, 0)
    with pytest.raises(RError, match="no CUDA device|no CPU path"):
        R.call("C_rfsi_near_obs", np.zeros((4, 2)), np.ones((3, 2)), np.arange(3.0), 2, True, 0)


def test_registration_table_as_r_would_load_it(R):
    import ctypes as C
    R.lib.R_init_rfsib200(None)                       # what R calls on dyn.load / library()
    R.lib.mock_registered.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int)]
    seen = {}
    i = 0
    name, n = C.c_char_p(), C.c_int()
    while R.lib.mock_registered(i, C.byref(name), C.byref(n)):
        seen[name.value.decode()] = n.value
        assert hasattr(R.lib, name.value.decode())
        i += 1
    assert {"C_rfsi_near_obs", "C_rfsi_predict", "C_rfsi_forest_new", "C_rfsi_pred_fused", "C_rfsi_pred_maps"} <= set(seen)
