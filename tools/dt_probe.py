"""near.obs host-call time (simulation.R's DT: 250 000 cells, n.obs 25, 500 stations, fresh result columns every call)
against the number of staging threads and the chunk size.  python tools/dt_probe.py"""
import os
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import rfsi_b200 as rb
from rfsi_b200 import synth

rng = np.random.default_rng(0)
loc = synth.lattice(500, 500)
obs = loc[rng.choice(250000, 500, replace=False)]
odf = np.column_stack([obs, rng.normal(size=500)])
keep = []


def dt(fresh):
    ts = []
    for _ in range(7):
        t0 = time.perf_counter(); r = rb.near_obs(loc, odf, zcol=2, n_obs=25); ts.append(time.perf_counter() - t0)
        if fresh:
            keep.append(r)          # never freed: the allocator cannot hand the same (already mapped) pages out again
            if len(keep) > 12: keep.pop(0)
    return 1e3 * min(ts[2:]), 1e3 * float(np.median(ts[2:]))


print("| threads | chunk | DT ms (min / median), columns reused | DT ms (min / median), fresh columns |")
print("|---|---|---|---|")
for th in (2, 4, 8, 16):
    for ch in (32768, 16384):
        os.environ["RFSI_B200_HOST_THREADS"] = str(th); os.environ["RFSI_B200_NEAR_CHUNK"] = str(ch)
        a = dt(False); b = dt(True)
        print(f"| {th} | {ch} | {a[0]:.2f} / {a[1]:.2f} | {b[0]:.2f} / {b[1]:.2f} |")
