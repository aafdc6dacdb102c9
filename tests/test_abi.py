"""The C-ABI library loads, exports every symbol include/rfsi_b200.h declares, validates
its inputs on the host, and refuses to compute without an sm_100 device (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import rfsi_b200 as rb
from rfsi_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "rfsi_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rfsi_[a-z0-9_]+)\s*\(", header))
    declared -= {"rfsi_fused_args"}
    lib = _lib.load()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert lib.rfsi_version() == 100


def test_struct_layout_matches_header():
    # a compile of the header with the host compiler gives the authoritative sizeof
    import subprocess, tempfile
    src = '#include "rfsi_b200.h"\n#include <stdio.h>\nint main(){printf("%zu\\n", sizeof(rfsi_fused_args_t));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        p = Path(td) / "s.c"
        p.write_text(src)
        subprocess.run(["gcc", "-I", str(ROOT / "include"), str(p), "-o", str(Path(td) / "s")], check=True)
        out = subprocess.run([str(Path(td) / "s")], check=True, capture_output=True, text=True).stdout
    assert int(out) == C.sizeof(_lib.FusedArgs)


def _stump():
    return dict(ntrees=1, node_offsets=np.array([0, 3], dtype=np.int64), left=np.array([1, 0, 0], dtype=np.int32),
                right=np.array([2, 0, 0], dtype=np.int32), split_var=np.array([0, 0, 0], dtype=np.int32),
                split_val=np.array([0.5, 1.0, 2.0]))


def _create(fo, nfeat=1):
    lib = _lib.load()
    h = C.c_void_p()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.rfsi_forest_create(fo["ntrees"], p(fo["node_offsets"]), p(fo["left"]), p(fo["right"]), p(fo["split_var"]),
                                p(fo["split_val"]), None, None, nfeat, C.byref(h))
    return rc, h


def test_forest_create_validates_on_host():
    lib = _lib.load()
    rc, h = _create(_stump())
    assert rc == 0 and h.value
    assert lib.rfsi_forest_info(h, 0) == 1 and lib.rfsi_forest_info(h, 3) == 1 and lib.rfsi_forest_info(h, 6) == 1
    lib.rfsi_forest_destroy(h)
    bad = _stump(); bad["left"][0] = 7                      # child out of range
    assert _create(bad)[0] == _lib.RFSI_E_FOREST
    bad = _stump(); bad["split_var"][0] = 3                 # variable outside nfeat
    assert _create(bad)[0] == _lib.RFSI_E_FOREST
    bad = _stump(); bad["left"][1] = 1; bad["right"][1] = 2  # node 1 points at itself: cycle
    assert _create(bad)[0] == _lib.RFSI_E_FOREST
    assert b"tree 0" in lib.rfsi_last_error()


def test_no_cpu_fallback_without_device(has_gpu):
    if has_gpu:
        pytest.skip("a device is present; the refusal path is covered on the CPU box")
    loc = np.zeros((4, 2)); obs = np.ones((3, 3))
    with pytest.raises(rb.RfsiError) as e:
        rb.near_obs(loc, obs, zcol=2, n_obs=1)
    assert e.value.code == _lib.RFSI_E_NO_DEVICE
    from rfsi_b200.forest import FlatForest
    fo = _stump()
    model = rb.RangerForest(FlatForest(1, fo["node_offsets"], fo["left"], fo["right"], fo["split_var"], fo["split_val"], ["x"]))
    with pytest.raises(rb.RfsiError) as e:
        rb.predict(model, np.zeros((5, 1)))
    assert e.value.code == _lib.RFSI_E_NO_DEVICE
    with pytest.raises(rb.RfsiError) as e:
        rb.predict_fused(model, obs_xy=np.zeros((3, 2)), obs_z=np.zeros((1, 3)), n_obs=1, locations=loc,
                         covariates={"x": np.zeros(4)})
    assert e.value.code in (_lib.RFSI_E_NO_DEVICE, _lib.RFSI_E_ARG)


def test_product_never_imports_the_oracle():
    for p in (ROOT / "rfsi_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".c", ".h", ".R"):
            txt = p.read_text(errors="ignore")
            assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt \
                and "libcpubaseline" not in txt, f"{p} touches the oracle"


def test_fork_after_a_device_call_is_refused_with_its_own_code():
    """No GPU needed: the library remembers the pid of its first device call; a forked child is told so
    (RFSI_E_FORKED) instead of meeting the CUDA runtime's "initialization error" (tests/test_gpu_fork.py is the
    GPU twin with real work in the children)."""
    import subprocess
    import sys
    import textwrap
    code = textwrap.dedent('''
        import os, sys
        sys.path.insert(0, %r)
        from rfsi_b200 import _lib
        lib = _lib.load()
        pid = os.fork()
        if pid == 0:                          # forked BEFORE any device call: the child may make its own
            os._exit(0 if lib.rfsi_device_count() >= 0 else 1)
        assert os.waitpid(pid, 0)[1] == 0
        assert lib.rfsi_device_count() >= 0   # first device call of the parent
        pid = os.fork()
        if pid == 0:
            rc = lib.rfsi_device_count()
            ok = rc == _lib.RFSI_E_FORKED and b"fork()" in lib.rfsi_last_error()
            os._exit(0 if ok else 1)
        assert os.waitpid(pid, 0)[1] == 0
        print("OK")
    ''') % str(ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
