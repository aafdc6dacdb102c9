#!/usr/bin/env python
"""bench.py -- predicted pixel-days/s of the RFSI hot path (near.obs features + forest).

Workload (BASELINE.json configs[4], the one the metric is quoted on): 10^4 stations,
10^4 x 10^4 raster, n.obs = 20, 500 trees, F = 43 (3 daily covariates + 20 dist + 20 obs),
2 % of stations missing per day.  One step = one band of raster rows x D days through
rfsi_dev_predict_fused (geometry pass + fused near.obs->forest kernel), inputs resident in
HBM.  `e2e` is the same band through rfsi_predict_fused with pinned HOST buffers (H2D of
the covariates and D2H of the predictions inside the timed region).  Weak scaling: every
rank owns its own band of rows; no collective on the compute path.

python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "predicted_pixels_per_sec"
UNIT = "pixel-days/s"
N_TREES = 500
TRAIN_DAYS = 20          # 20 days x 10^4 stations = 2 x 10^5 training rows


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=200, help="raster rows (of 10^4 cells) per step per GPU")
    ap.add_argument("--days", type=int, default=8, help="days per step")
    ap.add_argument("--trees", type=int, default=N_TREES)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--qrf", action="store_true", help="also time mean + 3 quantiles (IQR maps) on 2 days of the band")
    ap.add_argument("--forest-cache", default=None, help="npz file to keep the grown bench forest between runs (setup only)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------
# workload construction (shared by both arms)
# ------------------------------------------------------------------------------------
def build_case(days):
    from rfsi_b200 import synth
    return synth.make_case("c5", ndays=max(days, TRAIN_DAYS))


def grow(case, near_obs_fn, trees, quantreg=False, cache=None):
    from rfsi_b200 import grow as grower, synth
    from rfsi_b200.forest import FlatForest
    if cache and Path(cache).exists():
        z = np.load(cache)
        if int(z["ntrees"]) == trees and bool(z["quantreg"]) >= bool(quantreg):
            log(f"[bench] forest from {cache}")
            return FlatForest(trees, z["node_offsets"], z["left"], z["right"], z["split_var"], z["split_val"],
                              case.feature_names, z["rnv_offsets"] if quantreg else None, z["rnv"] if quantreg else None)
    t0 = time.time()
    X, y = synth.training_table(case, near_obs_fn)
    t1 = time.time()
    flat = grower.grow_forest(X, y, case.feature_names, num_trees=trees, min_node_size=5, seed=42, quantreg=quantreg)
    log(f"[bench] training table {X.shape} in {t1 - t0:.1f}s; grew {trees} trees "
        f"({int(flat.node_offsets[-1])} nodes) in {time.time() - t1:.1f}s")
    if cache:
        np.savez(cache, ntrees=trees, quantreg=bool(quantreg), node_offsets=flat.node_offsets, left=flat.left, right=flat.right,
                 split_var=flat.split_var, split_val=flat.split_val,
                 rnv_offsets=flat.rnv_offsets if quantreg else np.zeros(0, np.int64),
                 rnv=flat.random_node_values if quantreg else np.zeros(0))
    return flat


def cov_params(case):
    """(k[ncov, ndays, 16, 2], phase[ncov, ndays, 16]) of synth.smooth_field for the daily covariates."""
    ks, phs = [], []
    seed = 42
    for i, _ in enumerate(case.cov_names):
        kd, pd_ = [], []
        for d in range(case.obs_z.shape[0]):
            rng = np.random.default_rng(seed + 100 + 17 * i + 1000 * d)
            kd.append(rng.normal(size=(16, 2)) / 400.0)
            pd_.append(rng.uniform(0, 2 * np.pi, size=16))
        ks.append(kd); phs.append(pd_)
    return np.asarray(ks), np.asarray(phs)


# ------------------------------------------------------------------------------------
# CPU arm: the oracle port (kd-tree near.obs + threaded ranger-style descent)
# ------------------------------------------------------------------------------------
def cpu_rate(case, flat, days, budget_s, threads):
    """Time the CPU port on a bounded sample of the same workload; returns (rate, sample text)."""
    import oracle
    fo = flat.as_dict()

    def run(npix):
        loc = case.locations(0, npix)
        cov = case.covariates(loc, ndays=1)
        t0 = time.perf_counter()
        oracle.predict_fused(fo, case.feature_names, loc_xy=loc, obs_xy=case.obs_xy, obs_z=case.obs_z[:1],
                             n_obs=case.n_obs, covariates={k: v[:1] for k, v in cov.items()}, kdtree=True,
                             threads=threads)
        return time.perf_counter() - t0

    probe = 4096
    dt = run(probe)
    npix = int(min(max(probe, budget_s / max(dt, 1e-3) * probe), 2_000_000))
    npix = max(1024, npix // 1024 * 1024)
    dt = run(npix)
    return npix / dt, f"{npix} pixels x 1 day of the c5 band (rows 0..), {dt:.1f}s on {threads} threads", npix, dt


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    oracle.build()
    threads = oracle.hardware_threads()
    case = build_case(args.days)
    flat = grow(case, lambda l, o, z, n: oracle.near_obs(l, o, z, n, kdtree=True, nthreads=threads)[:2], args.trees)
    per_step = max(2.0, min(args.cpu_seconds, 180.0 / max(1, args.steps + args.warmup)))
    rates, sample = [], ""
    for i in range(args.warmup + args.steps):
        rate, sample, npix, dt = cpu_rate(case, flat, 1, per_step, threads)
        if i >= args.warmup:
            rates.append((npix, dt))
    tot_px = sum(r[0] for r in rates); tot_t = sum(r[1] for r in rates)
    value = tot_px / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, case, flat, cpu=True),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_config(args, case, flat, cpu=False):
    return {
        "workload": "c5 scale-up: 10^4 stations, 10^4x10^4 raster (band of rows per step), n.obs=20, "
                    f"{args.trees} trees, F=43 (3 daily covariates + 20 dist + 20 obs), 2% stations missing/day",
        "pixels_per_step_per_gpu": args.rows * case.ncol,
        "days_per_step": args.days,
        "forest_nodes": int(flat.node_offsets[-1]),
        "cpu_port": "kd-tree near.obs + threaded ranger-style descent (oracle/cpu_baseline.cpp); R/ranger not installable here",
        **({"reference_step": "a bounded sample of this workload (see cpu_baseline.sample): 1 day of a band of rows"}
           if cpu else {}),
    }


# ------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def main_ours(args):
    import torch
    import rfsi_b200 as rb
    from rfsi_b200 import _lib
    from rfsi_b200._lib import FusedArgs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        log(f"[bench] WORLD_SIZE {world} != --gpus {args.gpus}; using WORLD_SIZE")
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    if rb.device_count() < 1:
        raise SystemExit("bench.py: no sm_100 device visible (there is no CPU fallback)")

    # ---- inputs ------------------------------------------------------------------
    case = build_case(args.days)
    D, S, n_obs = args.days, case.obs_xy.shape[0], case.n_obs

    def gpu_near_obs(loc, obs, z, n):
        df = rb.near_obs(loc, np.column_stack([obs, z]), zcol=2, n_obs=n, device=local)
        a = df.to_numpy()
        return a[:, :n], a[:, n:]

    if rank == 0:
        flat = grow(case, gpu_near_obs, args.trees, quantreg=args.qrf and world == 1, cache=args.forest_cache)
        parts = [flat.node_offsets, flat.left, flat.right, flat.split_var, flat.split_val]
    else:
        flat, parts = None, None
    if world > 1:
        # replicate the forest rank 0 -> peers over NVLink (setup only; no collective on the compute path)
        meta = torch.zeros(2, dtype=torch.int64, device=dev)
        if rank == 0:
            meta[0] = len(flat.node_offsets); meta[1] = len(flat.left)
        dist.broadcast(meta, 0)
        n_off, n_nodes = int(meta[0]), int(meta[1])
        dts = [torch.int64, torch.int32, torch.int32, torch.int32, torch.float64]
        lens = [n_off, n_nodes, n_nodes, n_nodes, n_nodes]
        got = []
        for i, (dt, ln) in enumerate(zip(dts, lens)):
            t = torch.from_numpy(np.ascontiguousarray(parts[i])).to(dev) if rank == 0 else torch.empty(ln, dtype=dt, device=dev)
            dist.broadcast(t, 0)
            got.append(t.cpu().numpy())
        if rank != 0:
            from rfsi_b200.forest import FlatForest
            flat = FlatForest(args.trees, got[0], got[1], got[2], got[3], got[4], case.feature_names)
    model = rb.RangerForest(flat)
    model.upload(local)
    info = model.info()

    ncol = case.ncol
    P = args.rows * ncol
    band_rows = case.nrow // world
    row0 = rank * band_rows
    nsteps_total = args.warmup + args.steps + 2
    if args.rows * nsteps_total > band_rows:
        log("[bench] band shorter than steps x rows; steps wrap inside the band")

    obs_x = torch.from_numpy(np.ascontiguousarray(case.obs_xy[:, 0])).to(dev)
    obs_y = torch.from_numpy(np.ascontiguousarray(case.obs_xy[:, 1])).to(dev)
    obs_z = torch.from_numpy(np.ascontiguousarray(case.obs_z[:D])).to(dev)
    # daily covariates of the band, generated on the device with synth.smooth_field's formula
    kk, ph = cov_params(case)
    p = torch.arange(P, device=dev, dtype=torch.float64)
    px = 0.5 + torch.remainder(p, ncol)
    py = 0.5 + row0 + torch.div(p, ncol, rounding_mode="floor")
    cov = torch.empty((3, D, P), dtype=torch.float64, device=dev)
    for c in range(3):
        for d in range(D):
            k = torch.from_numpy(kk[c, d]).to(dev); f = torch.from_numpy(ph[c, d]).to(dev)
            cov[c, d] = np.sqrt(2.0 / 16) * torch.cos(px[:, None] * k[:, 0] + py[:, None] * k[:, 1] + f).sum(dim=1)
    del p, px, py
    out_mean = torch.empty((D, P), dtype=torch.float64, device=dev)
    fmap = rb.feature_map(case.feature_names, case.cov_names, n_obs)
    probs = None

    def make_args(step, device_side=True, host=None):
        a = FusedArgs()
        a.struct_size = C.sizeof(FusedArgs)
        a.npix = P
        a.grid_x0, a.grid_y0, a.grid_dx, a.grid_dy = case.grid[:4]
        a.grid_ncol = ncol
        a.pix_begin = (row0 + (step * args.rows) % max(1, band_rows - args.rows + 1)) * ncol
        a.nobs, a.ndays, a.n_obs, a.rm_dupl = S, D, n_obs, 1
        a.ncov = 3
        a.feat_map = fmap.ctypes.data_as(C.POINTER(C.c_int32))
        if device_side:
            a.obs_x, a.obs_y, a.obs_z = obs_x.data_ptr(), obs_y.data_ptr(), obs_z.data_ptr()
            ptrs = (C.c_void_p * 3)(*[cov[c].data_ptr() for c in range(3)])
            a.out_mean = out_mean.data_ptr()
        else:
            a.obs_x, a.obs_y, a.obs_z = host["ox"].ctypes.data, host["oy"].ctypes.data, host["oz"].ctypes.data
            ptrs = (C.c_void_p * 3)(*[host["cov"][c].data_ptr() for c in range(3)])
            a.out_mean = host["out"].data_ptr()
        strides = (C.c_int64 * 3)(P, P, P)
        a.cov = C.cast(ptrs, C.POINTER(C.c_void_p))
        a.cov_day_stride = C.cast(strides, C.POINTER(C.c_int64))
        a._keep = (ptrs, strides)
        return a

    a0 = make_args(0)
    ws_bytes = lib.rfsi_dev_fused_workspace_bytes(model.handle, C.byref(a0))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    counters = torch.zeros(4, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step_dev(i, count=False):
        a = make_args(i)
        rc = lib.rfsi_dev_predict_fused(model.handle, C.byref(a), ws.data_ptr(), counters.data_ptr() if count else None,
                                        local, C.c_void_p(stream.cuda_stream))
        _lib.check(rc)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # one counted pass: exact number of internal nodes visited per pixel-day for these inputs
    step_dev(0, count=True)
    torch.cuda.synchronize(dev)
    cnt = counters.cpu().numpy()
    visits_per_pd = float(cnt[0]) / max(1, int(cnt[1]))
    fallback_pd = int(cnt[2])
    assert int(cnt[1]) == P * D, f"counted {cnt[1]} pixel-days, expected {P * D}"
    assert bool(torch.isfinite(out_mean).all()), "non-finite prediction"

    for i in range(args.warmup):
        step_dev(1 + i)
    barrier()
    sampler = ClockSampler(local if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local]))
    sampler.start()
    lib.rfsi_profile_enable(1)
    launches0 = rb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step_dev(1 + args.warmup + i)
    e1.record(stream)
    barrier()
    lib.rfsi_profile_enable(0)
    launches = rb.launch_count() - launches0
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    prof = {}
    for cls, name in ((0, "knn_geometry"), (2, "fused"), (3, "fallback")):
        tms, n = C.c_double(0), C.c_uint64(0)
        lib.rfsi_profile_read(cls, C.byref(tms), C.byref(n))
        prof[name] = (tms.value, int(n.value))
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * P * D * args.steps / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel (fused near.obs->forest) ---------------------
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    f_ext = 3
    b_rf = 8 * f_ext + 16 * (visits_per_pd + args.trees) + 8          # SURVEY.md 8(d), fused form
    # DRAM traffic of the same kernel from the committed ncu --set full capture (per pixel-day)
    traffic = None
    tpath = ROOT / "profiles" / "traffic.json"
    if tpath.exists():
        try:
            traffic = float(json.loads(tpath.read_text())["fused_kernel"]["dram_bytes_per_pixel_day"])
        except Exception:
            traffic = None
    fused_ms, fused_n = prof["fused"]
    units_per_launch = P * D * args.steps / max(1, fused_n)
    achieved = b_rf * units_per_launch / (fused_ms / max(1, fused_n) * 1e-3) / 1e9 if fused_ms > 0 else None
    image = {0: "wide 16-byte fp64 nodes", 1: "compact 8-byte rank nodes", 2: "blocked 4-byte rank nodes, 4 levels per 128-byte line",
             3: "sector image: 4-byte rank nodes, 3 levels + implicit children per 32-byte sector, leaf payload in the leaf's block"}
    image_code = int(info.get("image", 3))
    roofline = {
        "bound": "hbm", "kernel": f"fused_kernel<256,J> on the {image[image_code]} "
                                  "(per-day candidate compaction + forest traversal)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
        "traffic": (traffic * units_per_launch) if traffic else None,
        "traffic_source": "profiles/traffic.json (ncu --set full dram__bytes_read+write, bench shape), bytes per launch",
        "peak_source": peak_src,
        "algorithmic_bytes_per_pixel_day": b_rf, "mean_internal_nodes_per_tree": visits_per_pd / args.trees,
        "pixel_days_per_launch": units_per_launch, "avg_launch_ms": fused_ms / max(1, fused_n),
        "kernel_share_of_step": fused_ms / ms if ms > 0 else None,
        "geometry_pass_ms_per_step": prof["knn_geometry"][0] / max(1, args.steps),
        "forest_bytes": info["device_bytes"],
        "note": "achieved counts every node visit as a 16-B gather (SURVEY.md 8d: algorithmic bytes); a warp's 32 "
                "adjacent pixels share tree paths and the rank-quantised images move 4 or 8 B per node, so L1/L2 serve most "
                "visits: frac > 1 of the HBM roof, DRAM traffic ~14x below the algorithmic figure; the kernel is "
                "L1-data-path bound (profiles/r01_e_*)",
    }

    # ---- e2e through the host-buffer C ABI -------------------------------------------
    e2e = None
    if not args.no_e2e:
        host = {
            "ox": np.ascontiguousarray(case.obs_xy[:, 0]), "oy": np.ascontiguousarray(case.obs_xy[:, 1]),
            "oz": np.ascontiguousarray(case.obs_z[:D]),
            "cov": [cov[c].cpu().pin_memory() for c in range(3)],
            "out": torch.empty((D, P), dtype=torch.float64).pin_memory(),
        }

        def step_host(i):
            a = make_args(i, device_side=False, host=host)
            _lib.check(lib.rfsi_predict_fused(model.handle, C.byref(a), local))

        for i in range(max(1, min(args.warmup, 2))):
            step_host(1 + i)
        # host API result == device API result on the same band
        step_host(0); step_dev(0); torch.cuda.synchronize(dev)
        assert torch.equal(host["out"], out_mean.cpu()), "host-buffer path and device path disagree"
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step_host(1 + args.warmup + i)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e = {"value": world * P * D * args.steps / float(t_e.item()), "unit": UNIT,
               "h2d_bytes_per_step": 3 * D * P * 8 + D * S * 8 + 2 * S * 8, "d2h_bytes_per_step": D * P * 8,
               "ms_per_step": 1e3 * float(t_e.item()) / args.steps,
               "api": "rfsi_predict_fused (pinned host buffers in, pinned host buffer out)"}

    # ---- quantile forest (IQR maps), reported separately -------------------
    qrf = None
    if args.qrf and world == 1:
        Pq, Dq = P, min(D, 2)   # the whole band: 4 bytes of leaf-id scratch per (pixel-day, tree)
        out_q = torch.empty((3, Dq, Pq), dtype=torch.float64, device=dev)
        out_m = torch.empty((Dq, Pq), dtype=torch.float64, device=dev)
        probs_c = (C.c_double * 3)(0.25, 0.5, 0.75)

        def qargs(step):
            a = make_args(step)
            a.npix, a.ndays = Pq, Dq
            a.probs = C.cast(probs_c, C.POINTER(C.c_double)); a.nprobs = 3
            a.out_mean, a.out_q = out_m.data_ptr(), out_q.data_ptr()
            return a
        aq = qargs(0)
        wsq = torch.empty(lib.rfsi_dev_fused_workspace_bytes(model.handle, C.byref(aq)), dtype=torch.uint8, device=dev)
        for i in range(2):
            _lib.check(lib.rfsi_dev_predict_fused(model.handle, C.byref(qargs(i)), wsq.data_ptr(), None, local, None))
        torch.cuda.synchronize(dev)
        lib.rfsi_profile_enable(1)
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record(stream)
        for i in range(args.steps):
            _lib.check(lib.rfsi_dev_predict_fused(model.handle, C.byref(qargs(2 + i)), wsq.data_ptr(), None, local, None))
        q1.record(stream)
        torch.cuda.synchronize(dev)
        lib.rfsi_profile_enable(0)
        qms = q0.elapsed_time(q1)
        tq, nq = C.c_double(0), C.c_uint64(0)
        lib.rfsi_profile_read(4, C.byref(tq), C.byref(nq))
        tf, nf = C.c_double(0), C.c_uint64(0)
        lib.rfsi_profile_read(2, C.byref(tf), C.byref(nf))
        qrf = {"value": Pq * Dq * args.steps / (qms * 1e-3), "unit": UNIT, "what": "mean + quantiles (.25,.5,.75)",
               "pixels": Pq, "days": Dq, "quantile_kernel_ms_per_step": tq.value / args.steps,
               "fused_kernel_ms_per_step": tf.value / args.steps}
        del out_q, out_m, wsq

    # ---- CPU baseline beside it (rank 0, N=1 only) -------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle
        threads = oracle.hardware_threads()
        rate, sample, _, _ = cpu_rate(case, flat, 1, args.cpu_seconds, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}

    if rank == 0:
        cfg = workload_config(args, case, flat)
        cfg.update({"cache": "inputs larger than L2 every step (forest %.0f MB, covariates %.0f MB, candidate lists "
                             "%.0f MB; a different band of rows each step)" %
                             (info["device_bytes"] / 1e6, 3 * D * P * 8 / 1e6, 32 * 12 * P / 1e6),
                    "parallelism": f"{world} x row bands, forest + stations replicated, no collective",
                    "fallback_pixel_days_per_step": fallback_pd})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "roofline": roofline,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if qrf is not None:
            line["qrf"] = qrf
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def emit(line: dict):
    """The one JSON line, written to the process's ORIGINAL stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    a = parse()
    # Libraries (NCCL's version banner, for one) write to fd 1 behind Python's back: park the real
    # stdout, point fd 1 at stderr for the whole run, and emit the JSON line on the parked fd.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    sys.exit(main_reference(a) if a.impl == "reference" else main_ours(a))
