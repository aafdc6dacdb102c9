"""Test helper: rfsi_b200/R/src/r_shim.c linked against tests/tools/mock_libR.c, so that the .Call
entries of the R package run without R.  `call(name, *args)` is `.Call(name, ...)`: numpy arrays go
in as numeric / integer vectors or matrices (column-major, like R), lists as R lists, None as NULL;
what comes back is converted the other way.  An R error raises RError with R's message."""
import ctypes as C
import subprocess
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
INTSXP, REALSXP, STRSXP, VECSXP, EXTPTRSXP, NILSXP = 13, 14, 16, 19, 22, 0


class RError(RuntimeError):
    pass


class ExtPtr:
    def __init__(self, sexp):
        self.sexp = sexp


class MockR:
    def __init__(self):
        so = Path(tempfile.mkdtemp(prefix="rmock")) / "rfsib200.so"
        libdir = ROOT / "rfsi_b200"
        subprocess.run(["gcc", "-shared", "-fPIC", "-O1", "-Wall", "-I", str(ROOT / "include"),
                        "-I", str(ROOT / "rfsi_b200" / "R" / "src" / "stub"), "-o", str(so),
                        str(ROOT / "tests" / "tools" / "mock_libR.c"), str(ROOT / "rfsi_b200" / "R" / "src" / "r_shim.c"),
                        f"-L{libdir}", "-lrfsi_b200", f"-Wl,-rpath,{libdir}"], check=True)
        L = self.lib = C.CDLL(str(so))
        P = C.c_void_p
        L.mock_call.restype = P; L.mock_call.argtypes = [P, C.c_int, C.POINTER(P)]
        L.mock_real.restype = P; L.mock_real.argtypes = [P, C.c_ssize_t, C.c_int, C.c_int]
        L.mock_int.restype = P; L.mock_int.argtypes = [P, C.c_ssize_t, C.c_int, C.c_int]
        L.mock_list.restype = P; L.mock_list.argtypes = [C.POINTER(P), C.c_int]
        L.mock_nil.restype = P
        L.mock_elt.restype = P; L.mock_elt.argtypes = [P, C.c_int]
        L.mock_data.restype = P; L.mock_data.argtypes = [P]
        L.mock_length.restype = C.c_ssize_t; L.mock_length.argtypes = [P]
        for f in ("mock_type", "mock_nrow", "mock_ncol"):
            getattr(L, f).restype = C.c_int; getattr(L, f).argtypes = [P]
        L.mock_name.restype = C.c_char_p; L.mock_name.argtypes = [P, C.c_int]
        L.mock_last_error.restype = C.c_char_p
        L.mock_last_warning.restype = C.c_char_p
        L.mock_finalize.argtypes = [P]
        L.mock_class.restype = C.c_char_p; L.mock_class.argtypes = [P]
        L.mock_rownames.restype = P; L.mock_rownames.argtypes = [P]

    # ---- Python -> SEXP ----
    def sexp(self, x):
        L = self.lib
        if x is None:
            return L.mock_nil()
        if isinstance(x, ExtPtr):
            return x.sexp
        if isinstance(x, (list, tuple)):
            elts = (C.c_void_p * len(x))(*[self.sexp(e) for e in x])
            return L.mock_list(elts, len(x))
        if isinstance(x, (bool, int, np.integer)):
            x = np.array([int(x)], dtype=np.int32)
        if isinstance(x, float):
            x = np.array([x])
        a = np.asarray(x)
        if a.ndim == 2:                      # R matrices are column-major
            flat = np.ascontiguousarray(a.T).ravel()
            nrow, ncol = a.shape
        else:
            flat, nrow, ncol = np.ascontiguousarray(a).ravel(), 0, 0
        if a.dtype.kind in "iub":
            flat = flat.astype(np.int32)
            return L.mock_int(flat.ctypes.data_as(C.c_void_p), flat.size, nrow, ncol)
        flat = flat.astype(np.float64)
        return L.mock_real(flat.ctypes.data_as(C.c_void_p), flat.size, nrow, ncol)

    # ---- SEXP -> Python ----
    def value(self, s):
        L = self.lib
        t = L.mock_type(s)
        if t == NILSXP:
            return None
        if t == EXTPTRSXP:
            return ExtPtr(s)
        if t == VECSXP:
            n = L.mock_length(s)
            vals = [self.value(L.mock_elt(s, i)) for i in range(n)]
            names = [L.mock_name(s, i).decode() for i in range(n)]
            if L.mock_class(s) == b"data.frame":
                import pandas as pd
                rn = self.value(L.mock_rownames(s))       # compact form c(NA, -nrow)
                df = pd.DataFrame(dict(zip(names, vals)))
                assert rn[0] == np.iinfo(np.int32).min and -rn[1] == len(df), "row.names must be c(NA_integer_, -nrow)"
                return df
            return dict(zip(names, vals)) if all(names) else vals
        n = L.mock_length(s)
        dt = np.int32 if t == INTSXP else np.float64
        a = np.ctypeslib.as_array(C.cast(L.mock_data(s), C.POINTER(C.c_int32 if t == INTSXP else C.c_double)), shape=(n,)).copy() \
            if n else np.empty(0, dtype=dt)
        if L.mock_ncol(s) or L.mock_nrow(s):
            return a.reshape(L.mock_ncol(s), L.mock_nrow(s)).T
        return a

    def call(self, name, *args):
        L = self.lib
        fn = C.cast(getattr(L, name), C.c_void_p)
        sx = (C.c_void_p * max(1, len(args)))(*[self.sexp(a) for a in args])
        before = L.mock_protect_depth()
        out = L.mock_call(fn, len(args), sx)
        assert L.mock_protect_depth() == before, f"{name}: PROTECT / UNPROTECT unbalanced"
        if not out:
            raise RError(L.mock_last_error().decode())
        return self.value(out)

    def last_warning(self):
        return self.lib.mock_last_warning().decode()

    def warning_count(self):
        return self.lib.mock_warning_count()

    def gc(self):
        self.lib.mock_reset()
