#!/usr/bin/env bash
# r02_t: per-feature rank-search tables with about one bucket per threshold (cap RFSI_B200_RANK_BUCKETS = 2^16 / 2^18 / 2^20)
#        + the leaf-flag sector image, against the previous build (librfsi_b200_base.so: 4096 buckets, leaf blocks visited).
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02t.sh'
set -u
mkdir -p gpurun_out
FC=/tmp/rfsi_bench_forest.npz
python -m pytest tests/test_gpu_forest.py tests/test_gpu_fuzz.py tests/test_gpu_maps.py tests/test_gpu_fused.py tests/test_gpu_api.py tests/test_gpu_multi.py tests/test_gpu_fullsize.py tests/test_gpu_raster.py -m gpu -x -q > gpurun_out/t_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/t_suite.log
B="python bench.py --no-cpu --no-e2e --no-configs --steps 4 --warmup 3 --forest-cache $FC"
for cap in 262144 1048576 65536; do
  RFSI_B200_RANK_BUCKETS=$cap $B > gpurun_out/t_cap$cap.json 2> gpurun_out/t_cap$cap.err; echo "cap $cap rc=$?"
done
if [ -f rfsi_b200/librfsi_b200_base.so ]; then
  RFSI_B200_LIB=rfsi_b200/librfsi_b200_base.so $B > gpurun_out/t_base.json 2> gpurun_out/t_base.err; echo "base rc=$?"
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/t_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); q = j.get('qrf') or {}; r = j["roofline"]
        print(f"{f[11:-5]:12s} value {j['value']/1e6:7.1f} M  fused ms/step {r.get('avg_launch_ms', 0):.2f}  d {r.get('mean_internal_nodes_per_tree', 0):.2f}  forest {r.get('forest_bytes', 0)/1e6:.0f} MB  "
              f"qrf {q.get('value', 0)/1e6:.1f} M (fused {q.get('fused_kernel_ms_per_step', 0):.2f} + q-kernel {q.get('quantile_kernel_ms_per_step', 0):.2f} ms)")
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
tail -n 3 gpurun_out/t_suite.log
