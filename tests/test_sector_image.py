"""The sector forest image (rfsi_b200/csrc/sector.cuh) on the host: its builder and the very walk the
GPU runs are compiled for the CPU by tests/tools/sector_host.cpp and held to the oracle -- terminal
nodes, tree-order sums (bit-equal), internal-node counts -- for every J the kernels are instantiated
with, tree counts that are not a multiple of J, tree ranges (chunked launches), single-node trees and
features that sit exactly on thresholds.  The GPU twin of this file is tests/test_gpu_forest.py."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle
from rfsi_b200 import grow, synth
from rfsi_b200.forest import FlatForest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = tmp_path_factory.mktemp("sector") / "libsector_host.so"
    subprocess.run(["g++", "-O2", "-fPIC", "-std=c++17", "-shared", "-Wall", "-Wno-unknown-pragmas",
                    "-o", str(so), str(ROOT / "tests" / "tools" / "sector_host.cpp")], check=True)
    L = C.CDLL(str(so))
    L.sh_stats.restype = C.c_int64
    L.sh_stats.argtypes = [C.c_void_p, C.c_int]
    L.sh_build.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.POINTER(C.c_void_p)]
    L.sh_free.argtypes = [C.c_void_p]
    L.sh_set_keys.argtypes = [C.c_void_p, C.c_void_p]
    L.sh_walk.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _build(lib, flat, nfeat):
    h = C.c_void_p()
    off = np.ascontiguousarray(flat.node_offsets, dtype=np.int64)
    assert lib.sh_build(flat.ntrees, _p(off), _p(flat.left), _p(flat.right), _p(flat.split_var), _p(flat.split_val),
                        nfeat, C.byref(h)) == 0
    return h


def _walk(lib, h, Q, T, J, t_begin=0, t_end=None, total=None):
    Xt = np.ascontiguousarray(Q.T)
    node = np.full((len(Q), T), -1, dtype=np.int32)
    key = np.zeros((len(Q), T), dtype=np.uint32)
    total = np.zeros(len(Q)) if total is None else total
    visits = C.c_uint64(0)
    assert lib.sh_walk(h, _p(Xt), len(Q), len(Q), J, t_begin, T if t_end is None else t_end, _p(total), _p(node), _p(key),
                       C.byref(visits)) == 0
    return total, node, key, visits.value


def _depths(flat):
    out = []
    for t in range(flat.ntrees):
        a, b = flat.node_offsets[t], flat.node_offsets[t + 1]
        left, right = flat.left[a:b], flat.right[a:b]
        d = np.zeros(b - a, dtype=np.int64)
        for n in range(b - a):                      # ranger numbers children after their parent
            if left[n] or right[n]:
                d[left[n]] = d[right[n]] = d[n] + 1
        out.append(d)
    return out


def test_sector_walk_equals_the_oracle(lib):
    case = synth.make_case("c3", ndays=8)
    X, y = synth.training_table(case, lambda l, o, z, n: oracle.near_obs(l, o, z, n)[:2])
    flat = grow.grow_forest(X, y, case.feature_names, num_trees=60, min_node_size=3, seed=1)
    rng = np.random.default_rng(0)
    Q = X[rng.choice(len(X), 400, replace=len(X) < 400)].copy()
    Q[200:] += rng.normal(0, 0.01, size=Q[200:].shape) * np.abs(Q[200:])          # points that are not training rows
    Q[:50] = np.round(Q[:50], 1)                                                   # features exactly on thresholds
    Q[50:60, 0] = np.unique(flat.split_val)[:10]
    h = _build(lib, flat, len(case.feature_names))
    fo = flat.as_dict()
    want_node = oracle.forest_terminal_nodes(fo, Q)
    want_mean = oracle.forest_predict(fo, Q)
    dep = _depths(flat)
    d = np.stack([dep[t][want_node[:, t]] for t in range(flat.ntrees)], axis=1)
    for J in (1, 2, 4, 8, 12, 16):                  # 60 trees: J = 8 leaves a tail of 4, J = 16 a tail of 12
        total, node, _, visits = _walk(lib, h, Q, flat.ntrees, J)
        np.testing.assert_array_equal(node, want_node, err_msg=f"J={J}")
        np.testing.assert_array_equal(total / flat.ntrees, want_mean, err_msg=f"J={J}")        # tree order: bit-equal
        assert visits == d.sum(), f"J={J}: internal nodes counted"
    # tree ranges (one launch per forest chunk): the running sum is carried through
    total, node, _, _ = _walk(lib, h, Q, flat.ntrees, 8, 0, 13)
    total, node2, _, _ = _walk(lib, h, Q, flat.ntrees, 8, 13, 60, total)
    np.testing.assert_array_equal(total / flat.ntrees, want_mean)
    np.testing.assert_array_equal(np.where(node2 >= 0, node2, node), want_node)
    blocks, leaves, unused = lib.sh_stats(h, 0), lib.sh_stats(h, 1), lib.sh_stats(h, 2)
    nodes = int(flat.node_offsets[-1])
    assert leaves == sum(int((flat.left[flat.node_offsets[t]:flat.node_offsets[t + 1]] == 0).sum()) for t in range(flat.ntrees))
    print(f"sector image: {blocks} blocks ({32 * blocks / nodes:.1f} B per ranger node), {leaves} leaf blocks, {unused} unused child slots")
    lib.sh_free(h)


def test_leaf_keys_travel_with_the_leaf_block(lib):
    case = synth.make_case("c1", ndays=2)
    X, y = synth.training_table(case, lambda l, o, z, n: oracle.near_obs(l, o, z, n)[:2])
    flat = grow.grow_forest(X, y, case.feature_names, num_trees=9, min_node_size=5, seed=3)
    h = _build(lib, flat, len(case.feature_names))
    nleaf = lib.sh_stats(h, 1)
    keys = (np.arange(nleaf, dtype=np.uint32) * np.uint32(2654435761)) ^ np.uint32(0x5bd1e995)
    lib.sh_set_keys(h, _p(keys))
    Q = X[:64]
    _, node, key, _ = _walk(lib, h, Q, flat.ntrees, 4)
    # leaf numbers run left to right within each tree, trees one after the other: recompute them from ranger's arrays
    base = 0
    for t in range(flat.ntrees):
        a, b = flat.node_offsets[t], flat.node_offsets[t + 1]
        left, right = flat.left[a:b], flat.right[a:b]
        order, st = {}, [0]
        while st:
            n = st.pop()
            if left[n] == 0 and right[n] == 0:
                order[n] = base + len(order)
            else:
                st.append(right[n]); st.append(left[n])
        np.testing.assert_array_equal(key[:, t], keys[[order[n] for n in node[:, t]]])
        base += len(order)
    lib.sh_free(h)


def test_degenerate_forests(lib):
    # a single-node tree, a stump, a left-only chain of depth 7 (pads, and blocks with one live child)
    trees = [
        dict(left=[0], right=[0], var=[0], val=[5.0]),
        dict(left=[1, 0, 0], right=[2, 0, 0], var=[1, 0, 0], val=[0.5, -1.0, 1.0]),
    ]
    left, right, var, val = [], [], [], []
    for k in range(7):                               # node 2k internal: left = leaf 2k+1, right = 2k+2
        left += [2 * k + 1, 0]; right += [2 * k + 2, 0]; var += [k % 2, 0]; val += [6.5 - k, 10.0 + k]
    left += [0]; right += [0]; var += [0]; val += [99.0]
    trees.append(dict(left=left, right=right, var=var, val=val))
    off = np.cumsum([0] + [len(t["left"]) for t in trees]).astype(np.int64)
    cat = lambda k, dt: np.concatenate([np.asarray(t[k], dtype=dt) for t in trees])
    flat = FlatForest(len(trees), off, cat("left", np.int32), cat("right", np.int32), cat("var", np.int32),
                      cat("val", np.float64), ["a", "b"])
    h = _build(lib, flat, 2)
    g = np.array([-1.0, 0.0, 0.5, 0.5000001, 1.5, 2.5, 3.5, 4.5, 5.5, 6.5, 7.0])
    Q = np.array([(a, b) for a in g for b in g])
    fo = flat.as_dict()
    for J in (1, 2, 4, 8):
        total, node, _, _ = _walk(lib, h, Q, 3, J)
        np.testing.assert_array_equal(node, oracle.forest_terminal_nodes(fo, Q))
        np.testing.assert_array_equal(total / 3, oracle.forest_predict(fo, Q))
    lib.sh_free(h)
