"""ctypes binding of librfsi_grow.so: the host-side forest grower (rfsi_b200/csrc/grow.cpp).

Training is not on the accelerated path; the grower stands behind
rfsi() and gives benchmarks forests of the named shapes without R/ranger.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

from .forest import FlatForest

LIB_PATH = Path(__file__).resolve().parent / "librfsi_grow.so"
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.rfsi_grow_forest.restype = C.c_int
        _lib.rfsi_grow_forest.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_int32, C.c_uint64, C.c_int32, C.c_int32] + \
                                         [C.POINTER(C.c_void_p)] * 6
        _lib.rfsi_grow_free.restype = None
        _lib.rfsi_grow_free.argtypes = [C.c_void_p]
    return _lib


def _take(ptr, n, dtype):
    a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,)).copy()
    _load().rfsi_grow_free(ptr)
    return a


def grow_forest(X, y, feature_names: Sequence[str], num_trees: int = 250, mtry: Optional[int] = None,
                min_node_size: int = 5, max_depth: int = 0, seed: int = 42, num_threads: int = 0,
                quantreg: bool = False) -> FlatForest:
    """ranger(formula, data, num.trees, mtry, min.node.size, quantreg) stand-in. X is (n, F)."""
    X = np.asarray(X, dtype=np.float64)
    n, F = X.shape
    Xc = np.ascontiguousarray(X.T)                    # column-major n x F
    yc = np.ascontiguousarray(y, dtype=np.float64)
    if np.isnan(Xc).any() or np.isnan(yc).any():
        raise ValueError("Missing data in training table")   # ranger refuses NA
    outs = [C.c_void_p() for _ in range(6)]
    rc = _load().rfsi_grow_forest(Xc.ctypes.data, yc.ctypes.data, n, F, int(num_trees), int(mtry or 0),
                                  int(min_node_size), int(max_depth), int(seed), int(num_threads),
                                  1 if quantreg else 0, *[C.byref(o) for o in outs])
    if rc != 0:
        raise RuntimeError(f"rfsi_grow_forest failed: {rc}")
    off = _take(outs[0], num_trees + 1, np.int64)
    total = int(off[-1])
    left = _take(outs[1], total, np.int32)
    right = _take(outs[2], total, np.int32)
    var = _take(outs[3], total, np.int32)
    val = _take(outs[4], total, np.float64)
    rnv = _take(outs[5], total, np.float64) if quantreg else None
    return FlatForest(int(num_trees), off, left, right, var, val, list(feature_names),
                      rnv_offsets=off[:-1].copy() if quantreg else None, random_node_values=rnv)
