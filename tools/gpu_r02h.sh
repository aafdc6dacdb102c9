#!/usr/bin/env bash
# r02_h: the one-pass quantile selection (qrf_row_fast) against the re-bucketing selection it replaces, on one box:
# quantile parity tests, A/B of the bench's quantile record, one --set full capture of the quantile phase.
#   gpurun --timeout 1200 -- 'bash tools/gpu_r02h.sh'
set -u
mkdir -p gpurun_out
FC=/tmp/rfsi_bench_forest.npz
python -m pytest tests/test_gpu_forest.py tests/test_gpu_fuzz.py tests/test_gpu_maps.py tests/test_gpu_fused.py -m gpu -x -q > gpurun_out/h_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/h_suite.log
RFSI_B200_QRF_FAST=0 python -m pytest tests/test_gpu_forest.py -m gpu -x -q -k quantile > gpurun_out/h_suite_old.log 2>&1; echo "suite(old selection) rc=$?" | tee -a gpurun_out/h_suite_old.log
B="python bench.py --no-cpu --no-e2e --no-configs --steps 4 --warmup 3 --forest-cache $FC"
$B > gpurun_out/h_fast.json 2> gpurun_out/h_fast.err; echo "fast rc=$?"
RFSI_B200_QRF_FAST=0 $B > gpurun_out/h_old.json 2> gpurun_out/h_old.err; echo "old rc=$?"
P1="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-configs --forest-cache $FC"
RFSI_BENCH_PROFILE_QRF=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"qrf_direct|fused_kernel" -c 3 -o gpurun_out/h_qrf_phase -f $P1 > gpurun_out/h_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/h_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); r = j["roofline"]; q = j.get('qrf') or {}
        print(f"{f[11:-5]:10s} value {j['value']/1e6:7.1f} M  fused launch {r['avg_launch_ms']:.2f} ms  qrf {q.get('value', 0)/1e6:.1f} M "
              f"(fused {q.get('fused_kernel_ms_per_step', 0):.2f} + q-kernel {q.get('quantile_kernel_ms_per_step', 0):.2f} ms) parity {q.get('parity_sample')}")
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
tail -3 gpurun_out/h_suite.log gpurun_out/h_suite_old.log; ls -la gpurun_out
