#!/usr/bin/env bash
# A/B of experiment builds (tools/build_variant.sh) and of J on one box, same cached forest.
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_scan.sh "lib:J lib:J ..."'   (lib "-" = the product library)
set -u
mkdir -p gpurun_out
FC=/tmp/rfsi_bench_forest.npz
B="python bench.py --no-cpu --no-e2e --steps 4 --warmup 3 --forest-cache $FC"
$B --qrf > gpurun_out/scan_base.json 2> gpurun_out/scan_base.err; echo "base rc=$?"; python -m pytest tests/test_gpu_forest.py tests/test_gpu_fused.py tests/test_gpu_fuzz.py tests/test_gpu_maps.py -m gpu -x -q > gpurun_out/scan_suite.log 2>&1; echo "suite rc=$?"
for v in $1; do
  lib=${v%%:*}; J=${v##*:}
  L=rfsi_b200/librfsi_b200.so; [ "$lib" != "-" ] && L=rfsi_b200/librfsi_b200_$lib.so
  RFSI_B200_LIB=$L RFSI_B200_J=$J $B > gpurun_out/scan_${lib}_J$J.json 2> gpurun_out/scan_${lib}_J$J.err; echo "$v rc=$?"
done
if [ -n "${CHECK_LIB:-}" ]; then
  RFSI_B200_LIB=rfsi_b200/librfsi_b200_$CHECK_LIB.so python -m pytest tests/test_gpu_forest.py tests/test_gpu_fused.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/scan_check.log 2>&1; echo "parity of $CHECK_LIB rc=$?"
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/scan_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); r = j["roofline"]
        q = j.get('qrf') or {}; print(f"{f[16:-5]:16s} value {j['value']/1e6:7.1f} M   fused launch {r['avg_launch_ms']:.2f} ms   qrf {q.get('value', 0)/1e6:.1f} M q-kernel {q.get('quantile_kernel_ms_per_step', 0):.2f} ms")
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
