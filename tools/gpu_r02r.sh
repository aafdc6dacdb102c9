#!/usr/bin/env bash
# r02_r: (1) quantile placement from arrival numbers (v4) against the two-round counting sort (v3);
#        (2) what an L2-resident forest would buy the walk: forests of 8 / 16 / 24 trees (49 / 98 / 147 MB) and the 500-tree
#            forest walked in L2-sized slices (RFSI_B200_CHUNK_MB: one launch per slice, feature construction repeated).
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02r.sh'
set -u
mkdir -p gpurun_out
FC=/tmp/rfsi_bench_forest.npz
python -m pytest tests/test_gpu_forest.py tests/test_gpu_fuzz.py tests/test_gpu_maps.py tests/test_gpu_fused.py -m gpu -x -q > gpurun_out/r_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/r_suite.log
B="python bench.py --no-cpu --no-e2e --no-configs --steps 4 --warmup 3"
$B --forest-cache $FC > gpurun_out/r_v4.json 2> gpurun_out/r_v4.err; echo "v4 rc=$?"
if [ -f rfsi_b200/librfsi_b200_v3.so ]; then
  RFSI_B200_LIB=rfsi_b200/librfsi_b200_v3.so $B --forest-cache $FC > gpurun_out/r_v3.json 2> gpurun_out/r_v3.err; echo "v3 rc=$?"
fi
for mb in 48 96 192; do
  RFSI_B200_CHUNK_MB=$mb $B --no-qrf --forest-cache $FC > gpurun_out/r_chunk$mb.json 2> gpurun_out/r_chunk$mb.err; echo "chunk $mb rc=$?"
done
for t in 8 16 24 48; do
  $B --no-qrf --trees $t > gpurun_out/r_trees$t.json 2> gpurun_out/r_trees$t.err; echo "trees $t rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); q = j.get('qrf') or {}; r = j["roofline"]
        print(f"{f[11:-5]:10s} value {j['value']/1e6:7.1f} M  launches {j.get('gpu_launches')}  fused ms/step {r.get('avg_launch_ms', 0):.2f}  geometry {r.get('geometry_pass_ms_per_step', 0):.2f} ms  forest {r.get('forest_bytes', 0)/1e6:.0f} MB  "
              f"qrf {q.get('value', 0)/1e6:.1f} M (fused {q.get('fused_kernel_ms_per_step', 0):.2f} + q-kernel {q.get('quantile_kernel_ms_per_step', 0):.2f} ms)")
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
tail -n 3 gpurun_out/r_suite.log
