#!/usr/bin/env bash
# r02_u: rank caches of the fused prologue (distance ranks of the candidates once per pixel, observation ranks once per
#        launch) against searching per pixel-day (RFSI_B200_RANK_CACHE=0), same build, same box.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02u.sh'
set -u
mkdir -p gpurun_out
FC=/tmp/rfsi_bench_forest.npz
python -m pytest tests/test_gpu_forest.py tests/test_gpu_fuzz.py tests/test_gpu_maps.py tests/test_gpu_fused.py tests/test_gpu_api.py tests/test_gpu_multi.py tests/test_gpu_fullsize.py tests/test_gpu_raster.py tests/test_gpu_r_shim.py -m gpu -x -q > gpurun_out/u_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/u_suite.log
B="python bench.py --no-cpu --no-configs --steps 4 --warmup 3 --forest-cache $FC"
$B > gpurun_out/u_cache.json 2> gpurun_out/u_cache.err; echo "cache rc=$?"
RFSI_B200_RANK_CACHE=0 $B > gpurun_out/u_nocache.json 2> gpurun_out/u_nocache.err; echo "nocache rc=$?"
$B --days 16 > gpurun_out/u_cache_d16.json 2> gpurun_out/u_cache_d16.err; echo "cache 16 days rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/u_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); q = j.get('qrf') or {}; r = j["roofline"]; e = j.get("e2e") or {}
        print(f"{f[11:-5]:12s} value {j['value']/1e6:7.1f} M  e2e {e.get('value', 0)/1e6:7.1f} M  fused ms/step {r.get('avg_launch_ms', 0):.2f}  geometry {r.get('geometry_pass_ms_per_step', 0):.2f} ms  launches {j.get('gpu_launches')}  "
              f"qrf {q.get('value', 0)/1e6:.1f} M (fused {q.get('fused_kernel_ms_per_step', 0):.2f} + q-kernel {q.get('quantile_kernel_ms_per_step', 0):.2f} ms)")
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
tail -n 3 gpurun_out/u_suite.log
