#!/usr/bin/env bash
# r02_i: quantile selection after the peel rewrite + fp32 mode of near.obs, one box.
#   gpurun --timeout 1200 -- 'bash tools/gpu_r02i.sh'
set -u
mkdir -p gpurun_out
FC=/tmp/rfsi_bench_forest.npz
python -m pytest tests/test_gpu_forest.py tests/test_gpu_fuzz.py tests/test_gpu_maps.py tests/test_gpu_fused.py tests/test_gpu_knn.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/i_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/i_suite.log
python tools/fp32_mode_time.py > gpurun_out/i_fp32_prefilter1.md 2> gpurun_out/i_fp32.err; echo "fp32 (prefilter) rc=$?"
RFSI_B200_KNN_F32_PREFILTER=0 python tools/fp32_mode_time.py > gpurun_out/i_fp32_prefilter0.md 2>> gpurun_out/i_fp32.err; echo "fp32 (no prefilter) rc=$?"
B="python bench.py --no-cpu --no-e2e --no-configs --steps 4 --warmup 3 --forest-cache $FC"
$B > gpurun_out/i_fast.json 2> gpurun_out/i_fast.err; echo "fast rc=$?"
P1="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-configs --forest-cache $FC"
RFSI_BENCH_PROFILE_QRF=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"qrf_direct" -c 1 -o gpurun_out/i_qrf -f $P1 > gpurun_out/i_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/i_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); r = j["roofline"]; q = j.get('qrf') or {}
        print(f"{f[11:-5]:10s} value {j['value']/1e6:7.1f} M  qrf {q.get('value', 0)/1e6:.1f} M "
              f"(fused {q.get('fused_kernel_ms_per_step', 0):.2f} + q-kernel {q.get('quantile_kernel_ms_per_step', 0):.2f} ms) parity {q.get('parity_sample')}")
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
tail -n 3 gpurun_out/i_suite.log; cat gpurun_out/i_fp32_prefilter1.md gpurun_out/i_fp32_prefilter0.md
