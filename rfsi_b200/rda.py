"""Read R workspaces (.rda / .RData / .rds) without R, and pull ranger forests out of them.

The reference exchanges models as `save(rfsi_model, file="../models/RFSI.rda")`
(temp_croatia/1_models.R:1063) and `load(file=...)` (temp_croatia/2_predictions.R:59,
prcp_catalonia/5_prcp_prediction.R:233).  R is not installed next to this engine, so this
module parses R's XDR serialisation (format versions 2 and 3, gzip/bzip2/xz) directly:
SURVEY.md 8(f) rank 3.  `write_rda` writes the subset needed to round-trip model lists, so
the reader can be tested on files of the exact shape ranger produces (no real model blob
ships with the reference: .MISSING_LARGE_BLOBS).

Supported: NULL, symbols, pairlists/language objects, environments, closures (incl. byte
code), logical/integer/double/complex/character/raw vectors, lists, expression vectors, S4
objects, external pointers, attributes, reference table, long vectors, ALTREP
(compact_intseq, compact_realseq, deferred_string, wrap_*).
"""
from __future__ import annotations

import bz2
import gzip
import io
import lzma
import struct
from typing import Any, Dict, List, Optional

import numpy as np

NA_INTEGER = -2147483648

# SEXP types
(NILSXP, SYMSXP, LISTSXP, CLOSXP, ENVSXP, PROMSXP, LANGSXP, SPECIALSXP, BUILTINSXP, CHARSXP, LGLSXP) = range(11)
INTSXP, REALSXP, CPLXSXP, STRSXP, DOTSXP, ANYSXP, VECSXP, EXPRSXP, BCODESXP, EXTPTRSXP, WEAKREFSXP, RAWSXP, S4SXP = (
    13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25)
REFSXP, NILVALUE_SXP, GLOBALENV_SXP, UNBOUNDVALUE_SXP, MISSINGARG_SXP, BASENAMESPACE_SXP = 255, 254, 253, 252, 251, 250
NAMESPACESXP, PACKAGESXP, PERSISTSXP, CLASSREFSXP, GENERICREFSXP, BCREPDEF, BCREPREF = 249, 248, 247, 246, 245, 244, 243
EMPTYENV_SXP, BASEENV_SXP, ATTRLANGSXP, ATTRLISTSXP, ALTREP_SXP = 242, 241, 240, 239, 238


class RList(list):
    """An R list (VECSXP) or pairlist: a Python list with `attrs`; elements by name too."""

    def __init__(self, items=(), attrs: Optional[dict] = None):
        super().__init__(items)
        self.attrs = attrs or {}

    @property
    def names(self) -> Optional[List[str]]:
        n = self.attrs.get("names")
        return None if n is None else list(n)

    def __getitem__(self, key):
        if isinstance(key, str):
            names = self.names or []
            if key not in names:
                raise KeyError(key)
            return super().__getitem__(names.index(key))
        return super().__getitem__(key)

    def get(self, key, default=None):
        try:
            return self[key]
        except (KeyError, IndexError):
            return default

    def keys(self):
        return self.names or []


class RArray(np.ndarray):
    """A numeric / logical R vector with attributes (dim applied in column-major order)."""

    def __new__(cls, data, attrs=None):
        obj = np.asarray(data).view(cls)
        obj.attrs = attrs or {}
        return obj

    def __array_finalize__(self, obj):
        self.attrs = getattr(obj, "attrs", {})


class RStrings(list):
    def __init__(self, items=(), attrs=None):
        super().__init__(items)
        self.attrs = attrs or {}


class RObject:
    """Anything else (S4 instance, environment, closure, language): kind + fields + attrs."""

    def __init__(self, kind: str, attrs=None, **fields):
        self.kind = kind
        self.attrs = attrs or {}
        self.fields = fields

    def __repr__(self):
        return f"<R {self.kind} {list(self.fields)} attrs={list(self.attrs)}>"


class _Reader:
    def __init__(self, data: bytes):
        self.b = io.BytesIO(data)
        self.refs: List[Any] = []

    # -- primitives ---------------------------------------------------------------------
    def int(self) -> int:
        return struct.unpack(">i", self.b.read(4))[0]

    def length(self) -> int:
        n = self.int()
        if n == -1:
            hi, lo = struct.unpack(">II", self.b.read(8))
            return (hi << 32) | lo
        return n

    def array(self, dtype, n):
        raw = self.b.read(n * np.dtype(dtype).itemsize)
        return np.frombuffer(raw, dtype=np.dtype(dtype).newbyteorder(">"), count=n).astype(dtype)

    # -- items ---------------------------------------------------------------------------
    def header(self):
        magic = self.b.read(2)
        if magic != b"X\n":
            raise ValueError("only the XDR serialisation format is supported")
        version = self.int()
        self.int(); self.int()          # writer version, minimal reader version
        if version == 3:
            n = self.int()
            self.b.read(n)              # native encoding
        elif version != 2:
            raise ValueError(f"unsupported serialisation version {version}")

    def attributes(self) -> dict:
        pl = self.item()
        out = {}
        if isinstance(pl, RList):
            for tag, val in zip(pl.attrs.get("tags", []), pl):
                out[tag] = val
        return out

    def item(self):
        flags = self.int()
        typ = flags & 0xFF
        has_attr, has_tag = bool(flags & (1 << 9)), bool(flags & (1 << 10))
        if typ == NILVALUE_SXP:
            return None
        if typ in (GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP, BASENAMESPACE_SXP, MISSINGARG_SXP, UNBOUNDVALUE_SXP):
            return RObject({GLOBALENV_SXP: "globalenv", EMPTYENV_SXP: "emptyenv", BASEENV_SXP: "baseenv",
                            BASENAMESPACE_SXP: "basenamespace", MISSINGARG_SXP: "missingarg",
                            UNBOUNDVALUE_SXP: "unbound"}[typ])
        if typ == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.int()
            return self.refs[idx - 1]
        if typ == PERSISTSXP or typ == PACKAGESXP or typ == NAMESPACESXP:
            self.int()                                  # 0
            n = self.int()
            s = [self.item() for _ in range(n)]
            obj = RObject({PERSISTSXP: "persist", PACKAGESXP: "package", NAMESPACESXP: "namespace"}[typ], info=s)
            self.refs.append(obj)
            return obj
        if typ == SYMSXP:
            name = self.item()
            self.refs.append(name)
            return name
        if typ == CHARSXP:
            n = self.int()
            if n == -1:
                return None
            raw = self.b.read(n)
            try:
                return raw.decode("utf-8")
            except UnicodeDecodeError:
                return raw.decode("latin-1")
        if typ == ENVSXP:
            env = RObject("environment")
            self.refs.append(env)
            self.int()                                  # locked
            env.fields["enclos"] = self.item()
            env.fields["frame"] = self.item()
            env.fields["hashtab"] = self.item()
            attr = self.item()
            env.attrs = {} if attr is None else dict(zip(attr.attrs.get("tags", []), attr))
            return env
        if typ in (LISTSXP, LANGSXP, CLOSXP, PROMSXP, DOTSXP, ATTRLISTSXP, ATTRLANGSXP):
            return self.pairlist(typ, has_attr, has_tag)
        if typ == EXTPTRSXP:
            obj = RObject("externalptr")
            self.refs.append(obj)
            obj.fields["prot"] = self.item()
            obj.fields["tag"] = self.item()
            obj.attrs = self.attributes() if has_attr else {}
            return obj
        if typ == WEAKREFSXP:
            obj = RObject("weakref")
            self.refs.append(obj)
            obj.attrs = self.attributes() if has_attr else {}
            return obj
        if typ in (SPECIALSXP, BUILTINSXP):
            n = self.int()
            return RObject("builtin", name=self.b.read(n).decode())
        if typ == S4SXP:
            return RObject("S4", attrs=self.attributes() if has_attr else {})
        if typ == BCODESXP:
            nreps = self.int()
            reps = [None] * nreps
            code = self.bytecode(reps)
            attrs = self.attributes() if has_attr else {}
            return RObject("bytecode", attrs=attrs, code=code)
        if typ == ALTREP_SXP:
            info = self.item()
            state = self.item()
            attr = self.item()
            val = self.altrep(info, state)
            if isinstance(attr, RList) and hasattr(val, "attrs"):
                val.attrs.update(dict(zip(attr.attrs.get("tags", []), attr)))
                val = self.finish_vector(val)
            return val
        # vectors
        if typ in (LGLSXP, INTSXP):
            val = RArray(self.array(np.int32, self.length()))
        elif typ == REALSXP:
            val = RArray(self.array(np.float64, self.length()))
        elif typ == CPLXSXP:
            n = self.length()
            val = RArray(self.array(np.float64, 2 * n).view(np.complex128))
        elif typ == RAWSXP:
            n = self.length()
            val = RArray(np.frombuffer(self.b.read(n), dtype=np.uint8).copy())
        elif typ == STRSXP:
            val = RStrings([self.item() for _ in range(self.length())])
        elif typ in (VECSXP, EXPRSXP):
            val = RList([self.item() for _ in range(self.length())])
        else:
            raise NotImplementedError(f"R object type {typ} is not supported")
        if has_attr:
            val.attrs = self.attributes()
        return self.finish_vector(val)

    @staticmethod
    def finish_vector(val):
        dim = val.attrs.get("dim") if hasattr(val, "attrs") else None
        if isinstance(val, RArray) and dim is not None and val.ndim == 1 and int(np.prod(dim)) == val.shape[0]:
            attrs = val.attrs
            val = RArray(np.asarray(val).reshape(tuple(int(d) for d in dim), order="F"), attrs)
        return val

    def pairlist(self, typ, has_attr, has_tag):
        items, tags = [], []
        attrs_first = None
        kind = typ
        while True:
            attr = self.attributes() if has_attr else None
            if attrs_first is None:
                attrs_first = attr or {}
            tag = self.item() if has_tag else None
            items.append(self.item())
            tags.append(tag)
            # CDR: iterate while it is another pairlist cell of the same family
            flags = self.int()
            t = flags & 0xFF
            if t in (LISTSXP, LANGSXP, ATTRLISTSXP, ATTRLANGSXP) and kind not in (CLOSXP, PROMSXP):
                has_attr, has_tag = bool(flags & (1 << 9)), bool(flags & (1 << 10))
                continue
            self.b.seek(-4, io.SEEK_CUR)
            tail = self.item()
            if kind in (CLOSXP, PROMSXP):
                names = ("environment", "formals", "body") if kind == CLOSXP else ("env", "value", "expr")
                return RObject("closure" if kind == CLOSXP else "promise", attrs=attrs_first,
                               **{names[0]: tag, names[1]: items[0], names[2]: tail})
            if tail is not None:
                items.append(tail)
                tags.append(None)
            out = RList(items, dict(attrs_first))
            out.attrs["tags"] = tags
            if any(t is not None for t in tags):
                out.attrs["names"] = ["" if t is None else t for t in tags]
            out.attrs["pairlist"] = "language" if kind in (LANGSXP, ATTRLANGSXP) else "pairlist"
            return out

    def bytecode(self, reps):
        code = self.item()
        n = self.int()
        consts = []
        for _ in range(n):
            t = self.int()
            if t == BCODESXP:
                consts.append(self.bytecode(reps))
            elif t in (LANGSXP, LISTSXP, BCREPDEF, BCREPREF, ATTRLANGSXP, ATTRLISTSXP):
                consts.append(self.bclang(t, reps))
            else:
                consts.append(self.item())
        return RObject("bytecode1", code=code, consts=consts)

    def bclang(self, t, reps):
        if t == BCREPREF:
            return reps[self.int()]
        if t in (BCREPDEF, LANGSXP, LISTSXP, ATTRLANGSXP, ATTRLISTSXP):
            pos = -1
            if t == BCREPDEF:
                pos = self.int()
                t = self.int()
            cell = RObject("bclang")
            if pos >= 0:
                reps[pos] = cell
            if t in (ATTRLANGSXP, ATTRLISTSXP):
                cell.attrs = {"attr": self.item()}
            cell.fields["tag"] = self.item()
            cell.fields["car"] = self.bclang(self.int(), reps)
            cell.fields["cdr"] = self.bclang(self.int(), reps)
            return cell
        return self.item()

    @staticmethod
    def altrep(info, state):
        cls = info[0] if isinstance(info, RList) else None
        if cls == "compact_intseq":
            n, start, step = (float(v) for v in np.asarray(state)[:3])
            return RArray((start + step * np.arange(int(n))).astype(np.int32))
        if cls == "compact_realseq":
            n, start, step = (float(v) for v in np.asarray(state)[:3])
            return RArray(start + step * np.arange(int(n), dtype=np.float64))
        if cls == "deferred_string":
            arg = state[0] if isinstance(state, RList) else state
            vals = np.asarray(arg)
            if np.issubdtype(vals.dtype, np.integer):
                return RStrings([None if v == NA_INTEGER else str(int(v)) for v in vals])
            return RStrings([None if np.isnan(v) else (str(int(v)) if float(v).is_integer() else repr(float(v)))
                             for v in vals])
        if cls is not None and cls.startswith("wrap_"):
            return state[0] if isinstance(state, RList) else state
        raise NotImplementedError(f"ALTREP class {cls!r} is not supported")


def _decompress(raw: bytes) -> bytes:
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


def read_rda(path) -> Dict[str, Any]:
    """load(file): {variable name: value}."""
    data = _decompress(open(path, "rb").read())
    if data[:5] not in (b"RDX2\n", b"RDX3\n"):
        raise ValueError(f"{path}: not an R workspace (magic {data[:5]!r})")
    r = _Reader(data[5:])
    r.header()
    top = r.item()
    if top is None:
        return {}
    return {name: val for name, val in zip(top.attrs.get("names", []), top)}


def read_rds(path) -> Any:
    """readRDS(file)."""
    r = _Reader(_decompress(open(path, "rb").read()))
    r.header()
    return r.item()


# ------------------------------------------------------------------------------------------
# ranger objects
# ------------------------------------------------------------------------------------------
def _is_ranger(obj) -> bool:
    if not isinstance(obj, RList):
        return False
    cls = obj.attrs.get("class")
    return (cls is not None and "ranger" in list(cls)) or ("forest" in obj.keys() and "num.trees" in obj.keys())


def ranger_to_dict(obj: RList) -> dict:
    """An R ranger object -> the dict layout rfsi_b200.forest.flatten_ranger() takes."""
    fo = obj["forest"]

    def scalar(v):
        return None if v is None else np.asarray(v).ravel()[0]

    forest = {
        "num.trees": int(scalar(fo["num.trees"])),
        "child.nodeIDs": [(np.asarray(t[0]), np.asarray(t[1])) for t in fo["child.nodeIDs"]],
        "split.varIDs": [np.asarray(v) for v in fo["split.varIDs"]],
        "split.values": [np.asarray(v, dtype=np.float64) for v in fo["split.values"]],
        "independent.variable.names": [str(s) for s in fo["independent.variable.names"]],
    }
    if fo.get("is.ordered") is not None:
        forest["is.ordered"] = [bool(v) for v in np.asarray(fo["is.ordered"]).ravel()]
    if fo.get("dependent.varID") is not None:
        forest["dependent.varID"] = int(scalar(fo["dependent.varID"]))
    model = {"forest": forest, "num.trees": forest["num.trees"]}
    tt = obj.get("treetype") or fo.get("treetype")
    if tt is not None:
        model["treetype"] = str(list(tt)[0])
    rnv = obj.get("random.node.values")
    if rnv is not None:
        model["random.node.values"] = np.asarray(rnv, dtype=np.float64)
    return model


def load_ranger(path, name: Optional[str] = None) -> dict:
    """The ranger model stored in an .rda/.RData (by variable name, or the only one) or .rds file."""
    data = _decompress(open(path, "rb").read())
    if data[:5] in (b"RDX2\n", b"RDX3\n"):
        ws = read_rda(path)
        if name is not None:
            return ranger_to_dict(ws[name])
        found = [k for k, v in ws.items() if _is_ranger(v)]
        if len(found) != 1:
            raise ValueError(f"{path}: expected exactly one ranger object, found {found or 'none'}")
        return ranger_to_dict(ws[found[0]])
    obj = read_rds(path)
    if not _is_ranger(obj):
        raise ValueError(f"{path}: not a ranger object")
    return ranger_to_dict(obj)


# ------------------------------------------------------------------------------------------
# writer (the subset needed for model lists): lets tests build files of ranger's shape
# ------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self):
        self.b = io.BytesIO()

    def int(self, v):
        self.b.write(struct.pack(">i", int(v)))

    def charsxp(self, s: Optional[str]):
        if s is None:
            self.int(CHARSXP); self.int(-1)
            return
        raw = s.encode("utf-8")
        self.int(CHARSXP | (1 << 15)); self.int(len(raw)); self.b.write(raw)   # UTF-8 flag in gp bits

    def symbol(self, s: str):
        self.int(SYMSXP); self.charsxp(s)       # (no reference table: every symbol written in full)

    def attrs(self, attrs: dict):
        for k, v in attrs.items():
            self.int(LISTSXP | (1 << 10))
            self.symbol(k)
            self.item(v)
        self.int(NILVALUE_SXP)

    def item(self, v, extra_attrs: Optional[dict] = None):
        attrs = dict(extra_attrs or {})
        if v is None:
            self.int(NILVALUE_SXP)
            return
        if isinstance(v, dict):
            attrs = {"names": list(v.keys()), **attrs}
            v = list(v.values())
        if isinstance(v, (bool, np.bool_)):
            v = np.array([v])
        if isinstance(v, (int, np.integer)):
            v = np.array([v], dtype=np.int32)
        if isinstance(v, float):
            v = np.array([v], dtype=np.float64)
        if isinstance(v, str):
            v = [v]
        flag = (1 << 9) if attrs or (isinstance(v, np.ndarray) and v.ndim > 1) else 0
        if isinstance(v, np.ndarray):
            if v.ndim > 1:
                attrs = {"dim": np.array(v.shape, dtype=np.int32), **attrs}
                v = v.reshape(-1, order="F")
            if v.dtype == np.bool_:
                self.int(LGLSXP | flag); self.int(len(v)); self.b.write(v.astype(">i4").tobytes())
            elif np.issubdtype(v.dtype, np.integer):
                self.int(INTSXP | flag); self.int(len(v)); self.b.write(v.astype(">i4").tobytes())
            else:
                self.int(REALSXP | flag); self.int(len(v)); self.b.write(v.astype(">f8").tobytes())
        elif isinstance(v, (list, tuple)) and all(isinstance(s, str) or s is None for s in v) and len(v) > 0:
            self.int(STRSXP | flag); self.int(len(v))
            for s in v:
                self.charsxp(s)
        elif isinstance(v, (list, tuple)):
            self.int(VECSXP | flag); self.int(len(v))
            for e in v:
                self.item(e)
        else:
            raise TypeError(f"cannot serialise {type(v)}")
        if flag:
            self.attrs(attrs)


def write_rda(path, objects: Dict[str, Any], version: int = 3, compress: str = "gzip", classes: Optional[dict] = None):
    """save(..., file=path) for dicts (named lists), lists, numpy vectors/matrices and strings.
    classes: {variable name: class attribute} for the top-level objects."""
    w = _Writer()
    w.b.write(b"X\n")
    w.int(version); w.int(0x040300); w.int(0x030500 if version == 3 else 0x020300)
    if version == 3:
        w.int(5); w.b.write(b"UTF-8")
    for name, val in objects.items():
        w.int(LISTSXP | (1 << 10))
        w.symbol(name)
        cls = (classes or {}).get(name)
        w.item(val, {"class": [cls] if isinstance(cls, str) else cls} if cls else None)
    w.int(NILVALUE_SXP)
    payload = (b"RDX3\n" if version == 3 else b"RDX2\n") + w.b.getvalue()
    if compress == "gzip":
        payload = gzip.compress(payload)
    elif compress == "bzip2":
        payload = bz2.compress(payload)
    elif compress == "xz":
        payload = lzma.compress(payload)
    with open(path, "wb") as f:
        f.write(payload)
