#!/usr/bin/env bash
# r02_v: the build with rank caches + leaf flags + per-feature rank tables: GPU tests, bench, launch list, ncu --set full of the fused kernel.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02v.sh'
set -u
mkdir -p gpurun_out
FC=/tmp/rfsi_bench_forest.npz
python -m pytest tests/test_gpu_forest.py tests/test_gpu_fuzz.py tests/test_gpu_maps.py tests/test_gpu_fused.py tests/test_gpu_api.py tests/test_gpu_multi.py tests/test_gpu_fullsize.py tests/test_gpu_raster.py tests/test_gpu_r_shim.py -m gpu -x -q > gpurun_out/v_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/v_suite.log
RFSI_B200_RANK_CACHE=0 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fuzz.py tests/test_gpu_maps.py -m gpu -x -q > gpurun_out/v_suite_nocache.log 2>&1; echo "suite (no rank cache) rc=$?" | tee -a gpurun_out/v_suite_nocache.log
B="python bench.py --no-cpu --no-configs --steps 4 --warmup 3 --forest-cache $FC"
$B > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench rc=$?"
P="python bench.py --steps 2 --warmup 1 --rows 200 --no-cpu --no-e2e --no-configs --forest-cache $FC"
$P > gpurun_out/v_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v_launches.csv $P > gpurun_out/v_ncu_launches.log 2>&1; echo "launch list rc=$?"
P1="python bench.py --steps 1 --warmup 1 --rows 200 --no-cpu --no-e2e --no-configs --no-qrf --forest-cache $FC"
$P1 > gpurun_out/v_plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 1 -o gpurun_out/v_fused -f $P1 > gpurun_out/v_ncu_fused.log 2>&1; echo "ncu fused rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/v_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); q = j.get('qrf') or {}; r = j["roofline"]; e = j.get("e2e") or {}
        print(f"{f[11:-5]:12s} value {j['value']/1e6:7.1f} M  e2e {e.get('value', 0)/1e6:7.1f} M  fused ms/step {r.get('avg_launch_ms', 0):.2f}  geometry {r.get('geometry_pass_ms_per_step', 0):.2f} ms  launches {j.get('gpu_launches')}  "
              f"qrf {q.get('value', 0)/1e6:.1f} M (fused {q.get('fused_kernel_ms_per_step', 0):.2f} + q-kernel {q.get('quantile_kernel_ms_per_step', 0):.2f} ms)")
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
tail -n 2 gpurun_out/v_suite.log gpurun_out/v_suite_nocache.log
