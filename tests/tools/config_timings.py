#!/usr/bin/env python
"""DT / PT timings of the reference's own configs through the R-call-shaped host API
(host buffers in, host buffers out -- what simulation.R:190-215 times as DT and PT).

python tests/tools/config_timings.py > profiles/r01_config_timings.md   (on a B200)
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import rfsi_b200 as rb  # noqa: E402
from rfsi_b200 import grow, synth  # noqa: E402


def best(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts)


def gpu_no(l, o, z, n):
    a = rb.near_obs(l, np.column_stack([o, z]), zcol=2, n_obs=n).to_numpy()
    return a[:, :n], a[:, n:]


def main():
    print("# rfsi_b200 on the reference's configs (B200, host-buffer API, best of 3 after 1 warm-up)\n")
    print("Reference numbers: simulation/accuracy_results.rda (unknown CPU, contended), BASELINE.md section 1.\n")
    print("## simulation.R as scripted: 500x500 grid (250 000 locations), n.obs = 25, 250 trees, F = 50\n")
    print("| stations | DT near.obs (s) | ref DT (s) | PT predict (s) | ref PT (s) | DT+PT px/s | ref px/s |")
    print("|---|---|---|---|---|---|---|")
    ref_dt = {100: 1.404, 200: 1.482, 500: 1.615, 1000: 1.646, 2000: 1.691, 5000: 1.755}
    ref_pt = {100: 2.930, 200: 3.370, 500: 4.053, 1000: 4.741, 2000: 5.596, 5000: 6.825}
    rng = np.random.default_rng(42)
    loc = synth.lattice(500, 500)
    for S in (100, 200, 500, 1000, 2000, 5000):
        obs = loc[rng.choice(250000, S, replace=False)]
        z = 20 + np.sqrt(10) * synth.smooth_field(obs, 1) + rng.normal(0, np.sqrt(2.5), S)
        odf = np.column_stack([obs, z])
        tr = rb.near_obs(obs, odf, zcol=2, n_obs=25)
        flat = grow.grow_forest(tr.to_numpy(), z, list(tr.columns), num_trees=250)
        model = rb.RangerForest(flat)
        df = rb.near_obs(loc, odf, zcol=2, n_obs=25)
        X = df.to_numpy()
        dt = best(lambda: rb.near_obs(loc, odf, zcol=2, n_obs=25))
        pt = best(lambda: rb.predict(model, X))
        print(f"| {S} | {dt:.4f} | {ref_dt[S]:.3f} | {pt:.4f} | {ref_pt[S]:.3f} | {250000 / (dt + pt):.3g} | "
              f"{250000 / (ref_dt[S] + ref_pt[S]):.3g} |")
        model.close()

    print("\n## sim_500.R: 500 stations, near.obs once with n.obs = 40\n")
    case = synth.make_case("c2")
    odf = np.column_stack([case.obs_xy, case.obs_z[0]])
    dt = best(lambda: rb.near_obs(case.locations(), odf, zcol=2, n_obs=40))
    print(f"near.obs(250 000 locations, 500 stations, n.obs = 40): {dt:.4f} s")

    print("\n## mapping loops (fused near.obs -> predict, mean + quantiles (.25,.5,.75) + Int16 products)\n")
    print("| config | pixels x days | forest | fused mean (s) | fused mean + quantiles (s) | pixel-days/s (mean) |")
    print("|---|---|---|---|---|---|")
    import oracle  # only to build training tables on the CPU here; not timed
    for name, train_days, days in (("c3", 366, 4), ("c4", 1095, 4)):
        # trained on every day of the record (~50 k / ~92 k rows), mapped on 4 days
        case = synth.make_case(name, ndays=train_days)
        X, y = synth.training_table(case, gpu_no)
        flat = grow.grow_forest(X, y, case.feature_names, num_trees=250, min_node_size=6, quantreg=True)
        model = rb.RangerForest(flat)
        cov = case.covariates(case.locations(), ndays=days)
        kw = dict(obs_xy=case.obs_xy, obs_z=case.obs_z[:days], n_obs=case.n_obs, grid=case.grid, npix=case.npix,
                  covariates=cov)
        tm = best(lambda: rb.predict_fused(model, **kw))
        tq = best(lambda: rb.predict_fused(model, quantiles=[.25, .5, .75], centi_int16=True, **kw))
        print(f"| {name} ({case.ncol}x{case.nrow}, {case.obs_xy.shape[0]} stations, n.obs={case.n_obs}, F={len(case.feature_names)}) | "
              f"{case.npix} x {days} | 250 trees, {int(flat.node_offsets[-1])} nodes | {tm:.4f} | {tq:.4f} | "
              f"{case.npix * days / tm:.3g} |")
        model.close()


if __name__ == "__main__":
    main()
