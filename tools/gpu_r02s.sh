#!/usr/bin/env bash
# r02_s: sector image with leaf flags in the third-level nodes (no visit to the leaf's block) against the previous build
#        (librfsi_b200_base.so), and the rank-search bucket count (4096 -> 32768 / 262144 buckets per feature; built on the base).
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02s.sh'
set -u
mkdir -p gpurun_out
FC=/tmp/rfsi_bench_forest.npz
python -m pytest tests/test_gpu_forest.py tests/test_gpu_fuzz.py tests/test_gpu_maps.py tests/test_gpu_fused.py tests/test_gpu_api.py tests/test_gpu_multi.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/s_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/s_suite.log
B="python bench.py --no-cpu --no-e2e --no-configs --steps 4 --warmup 3 --forest-cache $FC"
$B > gpurun_out/s_new.json 2> gpurun_out/s_new.err; echo "new rc=$?"
for v in base rb32k rb256k; do
  if [ -f rfsi_b200/librfsi_b200_$v.so ]; then
    RFSI_B200_LIB=rfsi_b200/librfsi_b200_$v.so $B > gpurun_out/s_$v.json 2> gpurun_out/s_$v.err; echo "$v rc=$?"
  fi
done
$B > gpurun_out/s_new2.json 2> gpurun_out/s_new2.err; echo "new (again) rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/s_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); q = j.get('qrf') or {}; r = j["roofline"]
        print(f"{f[11:-5]:10s} value {j['value']/1e6:7.1f} M  fused ms/step {r.get('avg_launch_ms', 0):.2f}  d {r.get('mean_internal_nodes_per_tree', 0):.2f}  forest {r.get('forest_bytes', 0)/1e6:.0f} MB  "
              f"qrf {q.get('value', 0)/1e6:.1f} M (fused {q.get('fused_kernel_ms_per_step', 0):.2f} + q-kernel {q.get('quantile_kernel_ms_per_step', 0):.2f} ms) parity {q.get('parity_sample')}")
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
tail -n 3 gpurun_out/s_suite.log
