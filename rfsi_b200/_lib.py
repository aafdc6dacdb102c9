"""ctypes binding of librfsi_b200.so (the C ABI declared in include/rfsi_b200.h).

The library is the product: there is no Python/NumPy compute path behind these
calls.  If the shared object is missing, or no sm_100 device is visible, calls fail
loudly (RfsiError) instead of falling back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "librfsi_b200.so"

RFSI_OK = 0
RFSI_W_TOO_FEW_OBS = 1
RFSI_E_CUDA = -1
RFSI_E_NO_DEVICE = -2
RFSI_E_ARG = -3
RFSI_E_NA_INPUT = -4
RFSI_E_OOM = -5
RFSI_E_FOREST = -6
RFSI_E_FORKED = -7

FEAT_DIST = 0x10000
FEAT_OBS = 0x20000


class RfsiError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"rfsi_b200 error {code}: {message}")
        self.code = code


class Raster(C.Structure):
    _fields_ = [
        ("data", C.c_void_p), ("dtype", C.c_int32), ("nlayers", C.c_int32), ("ncol", C.c_int64), ("nrow", C.c_int64),
        ("x_ul", C.c_double), ("y_ul", C.c_double), ("dx", C.c_double), ("dy", C.c_double),
        ("nodata", C.c_double), ("scale", C.c_double), ("offset", C.c_double),
    ]


RASTER_DIVIDE = 0x100
RASTER_DTYPES = {"float64": 0, "float32": 1, "int32": 2, "int16": 3, "uint16": 4, "uint8": 5}


class CovFix(C.Structure):
    _fields_ = [("target", C.c_int32), ("other", C.c_int32), ("delta", C.c_double)]


class FusedArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("loc_x", C.c_void_p),
        ("loc_y", C.c_void_p),
        ("npix", C.c_int64),
        ("grid_x0", C.c_double),
        ("grid_y0", C.c_double),
        ("grid_dx", C.c_double),
        ("grid_dy", C.c_double),
        ("grid_ncol", C.c_int64),
        ("pix_begin", C.c_int64),
        ("obs_x", C.c_void_p),
        ("obs_y", C.c_void_p),
        ("nobs", C.c_int32),
        ("obs_z", C.c_void_p),
        ("ndays", C.c_int32),
        ("n_obs", C.c_int32),
        ("rm_dupl", C.c_int32),
        ("ncov", C.c_int32),
        ("cov", C.POINTER(C.c_void_p)),
        ("cov_day_stride", C.POINTER(C.c_int64)),
        ("feat_map", C.POINTER(C.c_int32)),
        ("probs", C.POINTER(C.c_double)),
        ("nprobs", C.c_int32),
        ("out_mean", C.c_void_p),
        ("out_q", C.c_void_p),
        ("out_mean_i16", C.c_void_p),
        ("out_iqr_i16", C.c_void_p),
        ("pix_mask", C.c_void_p),
        ("cov_raster", C.POINTER(Raster)),
        ("cov_fix", C.POINTER(CovFix)),
        ("ncov_fix", C.c_int32),
        ("cov_loc_x", C.c_void_p),
        ("cov_loc_y", C.c_void_p),
    ]


# every symbol include/rfsi_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
_SIGNATURES = [
    ("rfsi_version", C.c_int, []),
    ("rfsi_last_error", C.c_char_p, []),
    ("rfsi_device_count", C.c_int, []),
    ("rfsi_host_alloc", C.c_int, [C.POINTER(_P), C.c_size_t]),
    ("rfsi_host_free", None, [_P]),
    ("rfsi_near_obs_f64", C.c_int, [_P, _P, C.c_int64, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, C.c_int]),
    ("rfsi_near_obs_cols_f64", C.c_int, [_P, _P, C.c_int64, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, C.c_int]),
    ("rfsi_near_obs_f32", C.c_int, [_P, _P, C.c_int64, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, C.c_int]),
    ("rfsi_dev_near_obs_f64", C.c_int,
     [_P, _P, C.c_int64, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, C.c_int, _P]),
    ("rfsi_near_obs_days_f64", C.c_int,
     [_P, _P, C.c_int64, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, C.c_int]),
    ("rfsi_raster_extract", C.c_int,
     [C.POINTER(Raster), C.c_int32, _P, _P, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64,
      C.c_int64, _P, C.c_int]),
    ("rfsi_forest_create", C.c_int, [C.c_int32, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.POINTER(_P)]),
    ("rfsi_forest_destroy", None, [_P]),
    ("rfsi_forest_upload", C.c_int, [_P, C.c_int]),
    ("rfsi_forest_info", C.c_int64, [_P, C.c_int]),
    ("rfsi_forest_predict", C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, C.c_int]),
    ("rfsi_forest_quantiles", C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, C.c_int32, _P, C.c_int]),
    ("rfsi_forest_terminal_nodes", C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, C.c_int]),
    ("rfsi_dev_forest_scratch_bytes", C.c_size_t, [_P, C.c_int64, C.c_int32]),
    ("rfsi_dev_forest_predict", C.c_int,
     [_P, _P, C.c_int64, C.c_int64, _P, _P, C.c_int32, _P, _P, _P, _P, _P, C.c_int, _P]),
    ("rfsi_predict_fused", C.c_int, [_P, C.POINTER(FusedArgs), C.c_int]),
    ("rfsi_predict_fused_multi", C.c_int, [_P, C.POINTER(FusedArgs), C.POINTER(C.c_int), C.c_int]),
    ("rfsi_dev_fused_workspace_bytes", C.c_size_t, [_P, C.POINTER(FusedArgs)]),
    ("rfsi_dev_predict_fused", C.c_int, [_P, C.POINTER(FusedArgs), _P, _P, C.c_int, _P]),
    ("rfsi_launch_count", C.c_uint64, []),
    ("rfsi_profile_enable", C.c_int, [C.c_int]),
    ("rfsi_profile_read", C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
]
EXPORTED_SYMBOLS = [s[0] for s in _SIGNATURES]

_lib = None


def load() -> C.CDLL:
    """Load the native library (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("RFSI_B200_LIB", LIB_PATH))
    if not path.exists():
        raise RfsiError(RFSI_E_NO_DEVICE,
                        f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    lib = C.CDLL(str(path))
    for name, restype, argtypes in _SIGNATURES:
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return load().rfsi_last_error().decode("utf-8", "replace")


def check(rc: int) -> int:
    """Raise on a negative RFSI_E_* code; return warnings (>0) to the caller."""
    if rc < 0:
        raise RfsiError(rc, last_error())
    return rc


def device_count() -> int:
    return int(load().rfsi_device_count())


def launch_count() -> int:
    return int(load().rfsi_launch_count())
