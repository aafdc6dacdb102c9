"""fp32 mode of near.obs against the fp64 call on the reference's stage-1-heavy shapes (simulation.R: 250 000 cells,
n.obs 25, 100..5000 stations; sim_500.R: n.obs 40): search-kernel time (profile class 5) and wall time of the host call.
Run twice -- RFSI_B200_KNN_F32_PREFILTER=1 (default) and =0 -- to separate what the fp32 distance arithmetic buys from
what the float buffers buy.  python tools/fp32_mode_time.py"""
import ctypes as C
import os
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import rfsi_b200 as rb
from rfsi_b200 import synth, _lib
lib = _lib.load()
rng = np.random.default_rng(0)
loc = synth.lattice(500, 500).astype(np.float32).astype(np.float64)


def timed(fn, reps=5):
    fn(); fn()
    lib.rfsi_profile_enable(1); fn(); lib.rfsi_profile_enable(0)
    t, n = C.c_double(0), C.c_uint64(0)
    lib.rfsi_profile_read(5, C.byref(t), C.byref(n))
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return t.value, best * 1e3


print(f"# RFSI_B200_KNN_F32_PREFILTER={os.environ.get('RFSI_B200_KNN_F32_PREFILTER', '1')}  (kernel ms = all search launches of one call; wall = best of 5 host calls)")
print("| stations | n.obs | fp64 kernel ms | fp32-mode kernel ms | fp64 call ms | fp32-mode call ms | same neighbours |")
print("|---|---|---|---|---|---|---|")
for S, n in ((100, 25), (200, 25), (500, 40), (1000, 25), (5000, 25), (20000, 25)):
    obs = loc[rng.choice(250000, S, replace=False)]
    odf = np.column_stack([obs, rng.normal(size=S).astype(np.float32)])
    k64, w64 = timed(lambda: rb.near_obs(loc, odf, zcol=2, n_obs=n))
    k32, w32 = timed(lambda: rb.near_obs(loc, odf, zcol=2, n_obs=n, precision="fp32"))
    _, i64 = rb.near_obs(loc, odf, zcol=2, n_obs=n, return_idx=True)
    _, i32 = rb.near_obs(loc, odf, zcol=2, n_obs=n, return_idx=True, precision="fp32")
    print(f"| {S} | {n} | {k64:.3f} | {k32:.3f} | {w64:.2f} | {w32:.2f} | {bool(np.array_equal(i64, i32))} |")
