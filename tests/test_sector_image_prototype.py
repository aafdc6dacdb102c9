"""PROTOTYPE check (tools/prototypes/sector_image.cpp): the 32-byte
"sector block" forest image -- three levels per sector, implicit children, leaf values inside the
leaf's own block -- is built from ranger's arrays and walked on the host; terminal nodes and
tree-order sums must equal the oracle's, and the sectors a walk touches are counted against what
the product's 128-byte blocks touch for the same paths.  Nothing here is on the product path."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle
from rfsi_b200 import grow, synth

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = tmp_path_factory.mktemp("proto") / "libsector_image.so"
    subprocess.run(["g++", "-O2", "-fPIC", "-std=c++17", "-shared", "-Wno-unknown-pragmas", "-o", str(so),
                    str(ROOT / "tools" / "prototypes" / "sector_image.cpp")], check=True)
    L = C.CDLL(str(so))
    L.sector_stats.restype = C.c_int64
    L.sector_stats.argtypes = [C.c_void_p, C.c_int]
    L.sector_walk.argtypes = [C.c_void_p] + [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sector_build.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.POINTER(C.c_void_p)]
    L.sector_free.argtypes = [C.c_void_p]
    L.sector_walk_lockstep.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    return L


def _depths(flat):
    """internal nodes above every node of every tree (node ids are ranger's, per tree)"""
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


def test_sector_image_walk_equals_the_oracle_and_touches_fewer_sectors(lib):
    case = synth.make_case("c3", ndays=8)
    X, y = synth.training_table(case, lambda l, o, z, n: oracle.near_obs(l, o, z, n)[:2])
    flat = grow.grow_forest(X, y, case.feature_names, num_trees=60, min_node_size=3, seed=1)
    rng = np.random.default_rng(0)
    Q = X[rng.choice(len(X), 400, replace=len(X) < 400)].copy()
    Q[200:] += rng.normal(0, 0.01, size=Q[200:].shape) * np.abs(Q[200:])          # and points that are not training rows
    Q[:50] = np.round(Q[:50], 1)                                    # some features exactly on thresholds
    Q[50:60, 0] = np.unique(flat.split_val)[:10]
    h = C.c_void_p()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    off = np.ascontiguousarray(flat.node_offsets, dtype=np.int64)
    assert lib.sector_build(flat.ntrees, p(off), p(flat.left), p(flat.right), p(flat.split_var), p(flat.split_val),
                            len(case.feature_names), C.byref(h)) == 0
    Xt = np.ascontiguousarray(Q.T)                                  # npix x F column-major
    node = np.empty((len(Q), flat.ntrees), dtype=np.int32)
    total = np.empty(len(Q))
    sectors, levels = C.c_int64(0), C.c_int64(0)
    lib.sector_walk(h, p(Xt), len(Q), len(Q), p(node), p(total), C.byref(sectors), C.byref(levels))
    fo = flat.as_dict()
    np.testing.assert_array_equal(node, oracle.forest_terminal_nodes(fo, Q))
    np.testing.assert_array_equal(total / flat.ntrees, oracle.forest_predict(fo, Q))        # summed in tree order: bit-equal
    # traffic for the same paths: d internal nodes above the leaf
    dep = _depths(flat)
    d = np.stack([dep[t][node[:, t]] for t in range(flat.ntrees)], axis=1)
    assert levels.value == d.sum()
    new = np.ceil(d / 3).astype(np.int64) + 1                       # one sector per 3 levels + the leaf's own (value included)
    assert sectors.value == new.sum()
    old = 3 * np.maximum(1, np.ceil(d / 4)).astype(np.int64) + 1    # 128-byte blocks: 3 sectors per 4 levels + the leaf-value gather
    blocks, nbytes = lib.sector_stats(h, 0), lib.sector_stats(h, 1)
    nodes = int(flat.node_offsets[-1])
    print(f"sector image: {blocks} blocks = {nbytes / 1e6:.1f} MB for {nodes} nodes ({nbytes / nodes:.1f} B/node; "
          f"{lib.sector_stats(h, 4)} child slots unreachable), mean depth {d.mean():.1f}: "
          f"{new.mean():.2f} sectors per tree walk vs {old.mean():.2f} with 128-byte blocks ({old.mean() / new.mean():.2f}x)")
    assert old.mean() / new.mean() > 1.8
    # the device loop itself (tools/prototypes/sector_walk.cuh compiled for the host): J chains in lock step, finished
    # chains parked on the idle block, tree counts that are not a multiple of J (60 trees: J = 8 leaves a tail of 4)
    for J in (1, 3, 4, 8):
        s2 = np.empty(len(Q)); n2 = np.empty_like(node)
        assert lib.sector_walk_lockstep(h, p(Xt), len(Q), len(Q), J, p(s2), p(n2)) == 0
        np.testing.assert_array_equal(n2, node, err_msg=f"J={J}")
        np.testing.assert_array_equal(s2, total, err_msg=f"J={J}")
    lib.sector_free(h)
