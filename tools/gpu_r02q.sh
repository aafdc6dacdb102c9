#!/usr/bin/env bash
# r02_q: quantile selection as a counting sort by bucket (qrf_row_fast v3) against the marking-pass form (previous build), one box.
#   gpurun --timeout 1200 -- 'bash tools/gpu_r02q.sh'
set -u
mkdir -p gpurun_out
FC=/tmp/rfsi_bench_forest.npz
python -m pytest tests/test_gpu_forest.py tests/test_gpu_fuzz.py tests/test_gpu_maps.py tests/test_gpu_fused.py tests/test_gpu_api.py tests/test_gpu_multi.py tests/test_gpu_r_shim.py -m gpu -x -q > gpurun_out/q_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/q_suite.log
B="python bench.py --no-cpu --no-e2e --no-configs --steps 4 --warmup 3 --forest-cache $FC"
$B > gpurun_out/q_new.json 2> gpurun_out/q_new.err; echo "new rc=$?"
if [ -f rfsi_b200/librfsi_b200_prev.so ]; then
  RFSI_B200_LIB=rfsi_b200/librfsi_b200_prev.so $B > gpurun_out/q_prev.json 2> gpurun_out/q_prev.err; echo "prev rc=$?"
fi
$B > gpurun_out/q_new2.json 2> gpurun_out/q_new2.err; echo "new (again) rc=$?"
P1="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-configs --forest-cache $FC"
RFSI_BENCH_PROFILE_QRF=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"qrf_direct" -c 1 -o gpurun_out/q_qrf -f $P1 > gpurun_out/q_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/q_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); q = j.get('qrf') or {}
        print(f"{f[11:-5]:10s} value {j['value']/1e6:7.1f} M  qrf {q.get('value', 0)/1e6:.1f} M "
              f"(fused {q.get('fused_kernel_ms_per_step', 0):.2f} + q-kernel {q.get('quantile_kernel_ms_per_step', 0):.2f} ms) parity {q.get('parity_sample')}")
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
tail -n 3 gpurun_out/q_suite.log
