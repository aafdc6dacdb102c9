#!/usr/bin/env python
"""Time the Croatian / Catalonian map loops (bench.py's c3 / c4 `configs` records) on their own, under whatever engine
switches the environment holds: python tools/maploop_probe.py [c3|c4] -> one line."""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import rfsi_b200 as rb  # noqa: E402
from rfsi_b200 import grow, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
if "--ballast" in sys.argv:       # what the map loops find when they run inside bench.py: torch loaded, tens of GB held on the device, pinned host memory
    import torch
    _ballast = [torch.empty(20 << 30, dtype=torch.uint8, device="cuda"), torch.empty(3 << 30, dtype=torch.uint8).pin_memory()]
    torch.cuda.synchronize()
train_days, days = (366, 4) if name == "c3" else (1095, 4)


def gpu_no(l, o, z, n):
    a = rb.near_obs(l, np.column_stack([o, z]), zcol=2, n_obs=n).to_numpy()
    return a[:, :n], a[:, n:]


case = synth.make_case(name, ndays=train_days)
Xt, yt = synth.training_table(case, gpu_no)
flat = grow.grow_forest(Xt, yt, case.feature_names, num_trees=250, min_node_size=6, quantreg=True)
model = rb.RangerForest(flat)
locs = case.locations()
cv = case.covariates(locs, ndays=days)
kw = dict(obs_xy=case.obs_xy, obs_z=case.obs_z[:days], n_obs=case.n_obs, grid=case.grid, npix=case.npix, covariates=cv)


def best(fn, reps=7):
    for _ in range(4):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return 1e3 * min(ts)


tm = best(lambda: rb.predict_fused(model, **kw))
tq = best(lambda: rb.predict_fused(model, quantiles=[.25, .5, .75], centi_int16=True, **kw))
sw = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("RFSI_B200_") and k != "RFSI_B200_TRACE")
print(f"{name} {sw or '(defaults)':60s} mean {tm:7.3f} ms   mean+quantiles+Int16 {tq:7.3f} ms   ({case.npix * days / tm / 1e3:.1f} M pixel-days/s)", flush=True)
if os.environ.get("RFSI_B200_TRACE"):
    rb.predict_fused(model, **kw)
