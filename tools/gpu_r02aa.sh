#!/usr/bin/env bash
# r02_aa: last small A/Bs on the final build: 4 CTAs of 64 registers per SM, two cached candidate shifts instead of three, fewer spare candidates.
set -u
mkdir -p gpurun_out
FC=/tmp/rfsi_bench_forest.npz
B="python bench.py --no-cpu --no-e2e --no-configs --no-qrf --steps 4 --warmup 3 --forest-cache $FC"
$B > gpurun_out/aa_base.json 2> gpurun_out/aa_base.err; echo "base rc=$?"
RFSI_B200_LIB=rfsi_b200/librfsi_b200_minb4.so $B > gpurun_out/aa_minb4.json 2> gpurun_out/aa_minb4.err; echo "minb4 rc=$?"
RFSI_B200_LIB=rfsi_b200/librfsi_b200_shifts2.so $B > gpurun_out/aa_shifts2.json 2> gpurun_out/aa_shifts2.err; echo "shifts2 rc=$?"
RFSI_B200_KC_EXTRA=4 $B > gpurun_out/aa_kcextra4.json 2> gpurun_out/aa_kcextra4.err; echo "kc_extra 4 rc=$?"
$B --days 16 > gpurun_out/aa_days16.json 2> gpurun_out/aa_days16.err; echo "16 days rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/aa_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); r = j["roofline"]
        print(f"{f[11:-5]:12s} value {j['value']/1e6:7.1f} M  fused ms/step {r.get('avg_launch_ms', 0):.2f}  geometry+fill {r.get('geometry_pass_ms_per_step', 0):.2f} ms  fallback pixel-days {j['config'].get('fallback_pixel_days_per_step')}")
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
