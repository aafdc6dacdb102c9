#!/usr/bin/env python
"""Warp-level memory traffic of the product's 128-byte block image and of the prototype sector image
(tools/prototypes/sector_image.cpp) for the SAME walks, on the bench workload's geometry: C5 stations,
8 x 4 pixel patches (one warp each), a forest grown like the bench forest.  CPU only.
python tests/tools/traffic_model.py [trees] [patches]"""
import ctypes as C
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle
from rfsi_b200 import grow, synth

trees = int(sys.argv[1]) if len(sys.argv) > 1 else 40
patches = int(sys.argv[2]) if len(sys.argv) > 2 else 48
so = Path(tempfile.mkdtemp()) / "libsector_image.so"
subprocess.run(["g++", "-O2", "-fPIC", "-std=c++17", "-shared", "-Wno-unknown-pragmas", "-o", str(so),
                str(ROOT / "tools" / "prototypes" / "sector_image.cpp")], check=True)
L = C.CDLL(str(so))
case = synth.make_case("c5", ndays=4)
X, y = synth.training_table(case, lambda l, o, z, n: oracle.near_obs(l, o, z, n, kdtree=True, nthreads=8)[:2])
t0 = time.time()
flat = grow.grow_forest(X, y, case.feature_names, num_trees=trees)
print(f"forest: {trees} trees, {int(flat.node_offsets[-1])} nodes, grown on {len(X)} rows in {time.time() - t0:.0f} s", file=sys.stderr)
rng = np.random.default_rng(0)
pix = []
for _ in range(patches):                                              # 8 x 4 patches at random places of the raster
    c0, r0 = int(rng.integers(0, case.ncol - 8)), int(rng.integers(0, case.nrow - 4))
    pix += [(r0 + r) * case.ncol + c0 + c for r in range(4) for c in range(8)]
pix = np.array(pix)
loc = np.column_stack([case.grid[0] + (pix % case.ncol) * case.grid[2], case.grid[1] + (pix // case.ncol) * case.grid[3]])
d = 0
ok = ~np.isnan(case.obs_z[d])
dist, obs, _, _ = oracle.near_obs(loc, case.obs_xy[ok], case.obs_z[d][ok], case.n_obs, kdtree=True, nthreads=8)
cols = []
for n in case.feature_names:
    if n.startswith("dist"):
        cols.append(dist[:, int(n[4:]) - 1])
    elif n.startswith("obs"):
        cols.append(obs[:, int(n[3:]) - 1])
    else:
        cols.append(case.cov_daily[n](loc, d))
Q = np.ascontiguousarray(np.stack(cols, axis=0))                       # F x npix = npix x F column-major
out = (C.c_int64 * 6)()
p = lambda a: a.ctypes.data_as(C.c_void_p)
off = np.ascontiguousarray(flat.node_offsets, dtype=np.int64)
L.warp_traffic(C.c_int(flat.ntrees), p(off), p(flat.left), p(flat.right), p(flat.split_var), p(flat.split_val), p(Q),
               C.c_int64(len(pix)), C.c_int64(len(pix)), out)
r4, s4, r3, s3, m4, m3 = [int(v) for v in out]
wt = patches * trees
print(f"{patches} warps (8x4 pixel patches) x {trees} trees, per warp and tree:")
print(f"  128-byte blocks (product): {r4 / wt:6.2f} load requests, {s4 / wt:6.2f} L1 sector lookups ({s4 / r4:.2f} per request), {m4 / wt:6.2f} sectors fetched")
print(f"  32-byte sector blocks    : {r3 / wt:6.2f} load requests, {s3 / wt:6.2f} L1 sector lookups ({s3 / r3:.2f} per request), {m3 / wt:6.2f} sectors fetched")
print(f"  ratio                    : {r4 / r3:.2f}x fewer requests, {s4 / s3:.2f}x fewer lookups, {m4 / m3:.2f}x fewer sectors fetched")
