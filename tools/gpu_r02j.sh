#!/usr/bin/env bash
# r02_j: quantile selection inside the fused kernel (fused.cuh, SELR) against the two-kernel form, one box.
#   gpurun --timeout 1200 -- 'bash tools/gpu_r02j.sh'
set -u
mkdir -p gpurun_out
FC=/tmp/rfsi_bench_forest.npz
python -m pytest tests/test_gpu_fused.py tests/test_gpu_maps.py tests/test_gpu_fuzz.py tests/test_gpu_forest.py tests/test_gpu_multi.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/j_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/j_suite.log
B="python bench.py --no-cpu --no-e2e --no-configs --steps 4 --warmup 3 --forest-cache $FC"
$B > gpurun_out/j_sel.json 2> gpurun_out/j_sel.err; echo "sel rc=$?"
RFSI_B200_QRF_FUSED=0 $B > gpurun_out/j_two.json 2> gpurun_out/j_two.err; echo "two-kernel rc=$?"
P1="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-configs --forest-cache $FC"
RFSI_BENCH_PROFILE_QRF=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"qrf_direct|fused_kernel<256" -c 2 -o gpurun_out/j_sel -f $P1 > gpurun_out/j_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/j_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); q = j.get('qrf') or {}
        print(f"{f[11:-5]:10s} value {j['value']/1e6:7.1f} M  qrf {q.get('value', 0)/1e6:.1f} M "
              f"(fused {q.get('fused_kernel_ms_per_step', 0):.2f} + q-kernel {q.get('quantile_kernel_ms_per_step', 0):.2f} ms)")
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
tail -n 5 gpurun_out/j_suite.log
