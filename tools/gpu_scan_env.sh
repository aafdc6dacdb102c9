#!/usr/bin/env bash
# A/B of environment switches on one box, same cached forest.
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_scan_env.sh "REFILL=0" "REFILL=1 J=12" ...'   (names without the RFSI_B200_ prefix)
set -u
mkdir -p gpurun_out
FC=/tmp/rfsi_bench_forest.npz
B="python bench.py --no-cpu --no-e2e --no-configs --steps 3 --warmup 2 --rows 400 --forest-cache $FC"
for v in "$@"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $(for kv in $v; do echo "RFSI_B200_$kv"; done) $B > gpurun_out/env_$tag.json 2> gpurun_out/env_$tag.err; echo "$v rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/env_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); r = j["roofline"]; q = j.get("qrf") or {}
        print(f"{f[15:-5]:24s} value {j['value']/1e6:7.1f} M   fused launch {r['avg_launch_ms']:.2f} ms   qrf {q.get('value', 0)/1e6:.1f} M "
              f"(fused {q.get('fused_kernel_ms_per_step', 0):.2f} + q-kernel {q.get('quantile_kernel_ms_per_step', 0):.2f} ms)")
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
