#!/usr/bin/env bash
# r02_ac: the final build (block address as LEA on a register base) -- whole GPU suite, smoke, the default bench line, ncu --set full of the
#         fused kernel for the roofline record's stamp, launch list.
set -u
mkdir -p gpurun_out
FC=/tmp/rfsi_bench_forest.npz
python -m pytest tests -m gpu -q > gpurun_out/ac_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/ac_suite.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ac_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/ac_smoke.log
python bench.py --forest-cache $FC > gpurun_out/ac_bench_n1.json 2> gpurun_out/ac_bench_n1.err; echo "bench rc=$?"
P1="python bench.py --steps 1 --warmup 1 --rows 200 --no-cpu --no-e2e --no-configs --no-qrf --forest-cache $FC"
$P1 > gpurun_out/ac_plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 1 -o gpurun_out/ac_fused -f $P1 > gpurun_out/ac_ncu_fused.log 2>&1; echo "ncu fused rc=$?"
P="python bench.py --steps 2 --warmup 1 --rows 200 --no-cpu --no-e2e --no-configs --forest-cache $FC"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ac_launches.csv $P > gpurun_out/ac_ncu_launches.log 2>&1; echo "launch list rc=$?"
tail -n 3 gpurun_out/ac_suite.log; tail -n 2 gpurun_out/ac_smoke.log
python - <<'PY'
import json
j = json.loads(open("gpurun_out/ac_bench_n1.json").read().strip().splitlines()[-1])
r = j["roofline"]; q = j.get("qrf") or {}; e = j.get("e2e") or {}; c = j.get("cpu_baseline") or {}
print(f"value {j['value']/1e6:.1f} M  e2e {e.get('value',0)/1e6:.1f} M  qrf {q.get('value',0)/1e6:.1f} M  cpu {c.get('value',0)/1e6:.3f} M ({c.get('cores')} cores)  frac {r.get('frac')}  stale {r.get('ncu_stale')}  parity {j.get('parity_sample')} {q.get('parity_sample')}")
PY
