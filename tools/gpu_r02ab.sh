#!/usr/bin/env bash
# r02_ab: forms of the block address in the sector walk (SASS per iteration of 8 chains: product 231 = 116 ALU + 64 FMA-pipe;
#         addr2 = immediate block size + base in registers: 222 = 116 + 56; addr3 = run-time block size + base in registers: 230 = 108 + 72)
set -u
mkdir -p gpurun_out
FC=/tmp/rfsi_bench_forest.npz
B="python bench.py --no-cpu --no-e2e --no-configs --no-qrf --steps 4 --warmup 3 --forest-cache $FC"
$B > gpurun_out/ab_base.json 2> gpurun_out/ab_base.err; echo "base rc=$?"
for v in addr2 addr3; do
  RFSI_B200_LIB=rfsi_b200/librfsi_b200_$v.so $B > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; echo "$v rc=$?"
done
$B > gpurun_out/ab_base2.json 2> gpurun_out/ab_base2.err; echo "base again rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); r = j["roofline"]
        print(f"{f[11:-5]:10s} value {j['value']/1e6:7.2f} M  fused ms/step {r.get('avg_launch_ms', 0):.2f}")
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
