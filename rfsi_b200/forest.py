"""Forest import: ranger model objects -> the flat arrays rfsi_forest_create() takes.

Mirrors the model object the reference saves and reloads
(temp_croatia/1_models.R:1056-1063 `ranger(..., quantreg=TRUE)` + `save(rfsi_model, ...)`;
temp_croatia/2_predictions.R:59 and prcp_catalonia/5_prcp_prediction.R:233 `load(...)`):
`$forest$child.nodeIDs`, `$forest$split.varIDs`, `$forest$split.values`,
`$forest$independent.variable.names`, `$random.node.values`.

Training is NOT part of the accelerated path.  `from_sklearn` exists because no trained
ranger forest ships with the reference (.MISSING_LARGE_BLOBS) and R is not available
here: it converts a scikit-learn forest into ranger's layout so tests and benchmarks
have realistic forests.  scikit-learn is a forest *source* only, never an oracle (its own
predict() casts X to float32).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class FlatForest:
    """ranger layout, trees concatenated (the arguments of rfsi_forest_create)."""
    ntrees: int
    node_offsets: np.ndarray          # int64 [T+1]
    left: np.ndarray                  # int32, tree-local child ids, 0 = none
    right: np.ndarray
    split_var: np.ndarray             # int32, 0-based column of X
    split_val: np.ndarray             # float64, threshold or leaf prediction
    feature_names: list
    rnv_offsets: Optional[np.ndarray] = None      # int64 [T]
    random_node_values: Optional[np.ndarray] = None  # float64 flat

    @property
    def nfeat(self) -> int:
        return len(self.feature_names)

    def as_dict(self) -> dict:
        return dict(ntrees=self.ntrees, node_offsets=self.node_offsets, left=self.left, right=self.right,
                    split_var=self.split_var, split_val=self.split_val, rnv_offsets=self.rnv_offsets,
                    random_node_values=self.random_node_values)


def flatten_ranger(model: dict) -> FlatForest:
    """`model` mirrors the R list: model["forest"]["child.nodeIDs"][t] = (left, right), etc.

    Handles the old ranger convention where split.varIDs count the response column
    (`forest$dependent.varID` present): ids above it are shifted down by one
    (SURVEY.md Appendix A.2 item 2).
    """
    fo = model["forest"]
    T = int(fo["num.trees"])
    names = list(fo["independent.variable.names"])
    if model.get("treetype", "Regression") != "Regression":
        raise ValueError("only regression forests are supported (treetype == 'Regression')")
    is_ordered = fo.get("is.ordered")
    if is_ordered is not None and not all(bool(v) for v in np.atleast_1d(is_ordered)):
        raise ValueError("unordered (factor) splits are not supported: all RFSI features are numeric")
    dep = fo.get("dependent.varID")
    lefts, rights, vars_, vals, offs = [], [], [], [], [0]
    for t in range(T):
        l = np.asarray(fo["child.nodeIDs"][t][0], dtype=np.int64)
        r = np.asarray(fo["child.nodeIDs"][t][1], dtype=np.int64)
        v = np.asarray(fo["split.varIDs"][t], dtype=np.int64)
        s = np.asarray(fo["split.values"][t], dtype=np.float64)
        if not (len(l) == len(r) == len(v) == len(s)):
            raise ValueError(f"tree {t}: child/split arrays differ in length")
        if dep is not None:
            v = np.where(v > int(dep), v - 1, v)
        lefts.append(l); rights.append(r); vars_.append(v); vals.append(s)
        offs.append(offs[-1] + len(l))
    rnv = model.get("random.node.values")
    rnv_off = rnv_flat = None
    if rnv is not None:
        rnv = np.asarray(rnv, dtype=np.float64)       # (max_nodes, T) like the R matrix
        if rnv.ndim != 2 or rnv.shape[1] != T:
            raise ValueError("random.node.values must be a (max nodes x num.trees) matrix")
        rnv_flat = np.ascontiguousarray(rnv.T).ravel()  # column t contiguous = R column-major
        rnv_off = (np.arange(T, dtype=np.int64) * rnv.shape[0])
    return FlatForest(
        ntrees=T,
        node_offsets=np.asarray(offs, dtype=np.int64),
        left=np.concatenate(lefts).astype(np.int32),
        right=np.concatenate(rights).astype(np.int32),
        split_var=np.concatenate(vars_).astype(np.int32),
        split_val=np.concatenate(vals).astype(np.float64),
        feature_names=names,
        rnv_offsets=rnv_off,
        random_node_values=rnv_flat,
    )


def ranger_dict_from_flat(ff: FlatForest) -> dict:
    """The inverse of flatten_ranger: a dict shaped like R's ranger object."""
    off = ff.node_offsets
    T = ff.ntrees
    model = {
        "num.trees": T,
        "treetype": "Regression",
        "forest": {
            "num.trees": T,
            "child.nodeIDs": [(ff.left[off[t]:off[t + 1]].copy(), ff.right[off[t]:off[t + 1]].copy()) for t in range(T)],
            "split.varIDs": [ff.split_var[off[t]:off[t + 1]].copy() for t in range(T)],
            "split.values": [ff.split_val[off[t]:off[t + 1]].copy() for t in range(T)],
            "independent.variable.names": list(ff.feature_names),
            "is.ordered": [True] * ff.nfeat,
        },
    }
    if ff.random_node_values is not None:
        n = np.diff(off)
        m = int(n.max())
        rnv = np.full((m, T), np.nan)
        for t in range(T):
            rnv[: n[t], t] = ff.random_node_values[ff.rnv_offsets[t]: ff.rnv_offsets[t] + n[t]]
        model["random.node.values"] = rnv
    return model


def from_sklearn(rf, feature_names: Sequence[str], X_train=None, y_train=None, quantreg: bool = False,
                 seed: int = 0) -> FlatForest:
    """scikit-learn RandomForestRegressor / ExtraTreesRegressor -> ranger layout.

    quantreg=True also builds ranger's `random.node.values`: per tree, the training rows
    are dropped down the tree in a random order and each row's response overwrites its
    terminal node's cell, leaving one random response per terminal node (Appendix A.3.1).
    """
    lefts, rights, vars_, vals, offs = [], [], [], [], [0]
    rnv_parts, rnv_off = [], []
    rng = np.random.default_rng(seed)
    for est in rf.estimators_:
        tr = est.tree_
        leaf = tr.children_left < 0
        lefts.append(np.where(leaf, 0, tr.children_left).astype(np.int32))
        rights.append(np.where(leaf, 0, tr.children_right).astype(np.int32))
        vars_.append(np.where(leaf, 0, tr.feature).astype(np.int32))
        vals.append(np.where(leaf, tr.value[:, 0, 0], tr.threshold).astype(np.float64))
        if quantreg:
            if X_train is None or y_train is None:
                raise ValueError("quantreg=True needs X_train and y_train")
            term = est.apply(np.asarray(X_train, dtype=np.float32))
            col = np.full(tr.node_count, np.nan)
            order = rng.permutation(len(term))
            col[term[order]] = np.asarray(y_train, dtype=np.float64)[order]  # last write wins
            rnv_off.append(sum(len(p) for p in rnv_parts))
            rnv_parts.append(col)
        offs.append(offs[-1] + tr.node_count)
    return FlatForest(
        ntrees=len(rf.estimators_),
        node_offsets=np.asarray(offs, dtype=np.int64),
        left=np.concatenate(lefts), right=np.concatenate(rights),
        split_var=np.concatenate(vars_), split_val=np.concatenate(vals),
        feature_names=list(feature_names),
        rnv_offsets=np.asarray(rnv_off, dtype=np.int64) if quantreg else None,
        random_node_values=np.concatenate(rnv_parts) if quantreg else None,
    )


class RangerForest:
    """A ranger regression forest resident in the engine (rfsi_forest_t handle)."""

    def __init__(self, flat: FlatForest):
        self.flat = flat
        self.feature_names = list(flat.feature_names)
        lib = _lib.load()
        h = C.c_void_p()
        # keep the arrays alive and contiguous for the call
        arrs = [np.ascontiguousarray(flat.node_offsets, dtype=np.int64),
                np.ascontiguousarray(flat.left, dtype=np.int32),
                np.ascontiguousarray(flat.right, dtype=np.int32),
                np.ascontiguousarray(flat.split_var, dtype=np.int32),
                np.ascontiguousarray(flat.split_val, dtype=np.float64)]
        rnv_o = None if flat.rnv_offsets is None else np.ascontiguousarray(flat.rnv_offsets, dtype=np.int64)
        rnv = None if flat.random_node_values is None else np.ascontiguousarray(flat.random_node_values, dtype=np.float64)
        rc = lib.rfsi_forest_create(flat.ntrees, *[_ptr(a) for a in arrs], _ptr(rnv_o), _ptr(rnv), flat.nfeat,
                                    C.byref(h))
        _lib.check(rc)
        self._h = h

    @classmethod
    def from_ranger(cls, model: dict) -> "RangerForest":
        return cls(flatten_ranger(model))

    @classmethod
    def from_rda(cls, path, name: Optional[str] = None) -> "RangerForest":
        """load("../models/RFSI.rda") without R (temp_croatia/2_predictions.R:59): the ranger
        object saved in an .rda/.RData/.rds file, by variable name or the only one in the file."""
        from . import rda
        return cls(flatten_ranger(rda.load_ranger(path, name)))

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("forest was destroyed")
        return self._h

    @property
    def num_trees(self) -> int:
        return self.flat.ntrees

    @property
    def has_quantile_samples(self) -> bool:
        return self.flat.random_node_values is not None

    def info(self) -> dict:
        lib = _lib.load()
        keys = ["ntrees", "nfeat", "nodes", "max_depth", "has_samples", "device_bytes", "internal_nodes", "image",
                "unique_thresholds", "blocks"]          # image: 0 wide, 1 compact, 2 blocked
        return {k: int(lib.rfsi_forest_info(self.handle, i)) for i, k in enumerate(keys)}

    def upload(self, device: int = 0):
        _lib.check(_lib.load().rfsi_forest_upload(self.handle, device))

    def close(self):
        if getattr(self, "_h", None) is not None:
            _lib.load().rfsi_forest_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
