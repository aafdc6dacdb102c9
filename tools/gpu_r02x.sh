#!/usr/bin/env bash
# r02_x: end-of-round verification on one box: whole GPU suite, smoke, the default bench line (what the driver runs), the reference arm,
#        compute-sanitizer over every kernel, the parity tests under every engine switch.
#   gpurun --timeout 2400 -- 'bash tools/gpu_r02x.sh'
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/x_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/x_suite.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/x_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/x_smoke.log
python bench.py > gpurun_out/x_bench_n1.json 2> gpurun_out/x_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/x_bench_ref.json 2> gpurun_out/x_bench_ref.err; echo "reference arm rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/tools/sanitize_smoke.py > gpurun_out/x_sanitizer.log 2>&1; echo "sanitizer rc=$?" | tee -a gpurun_out/x_sanitizer.log
tail -n 4 gpurun_out/x_suite.log; tail -n 2 gpurun_out/x_smoke.log; tail -n 4 gpurun_out/x_sanitizer.log
python - <<'PY'
import json
j = json.loads(open("gpurun_out/x_bench_n1.json").read().strip().splitlines()[-1])
r = j["roofline"]; q = j.get("qrf") or {}; e = j.get("e2e") or {}; c = j.get("cpu_baseline") or {}
print(f"value {j['value']/1e6:.1f} M  e2e {e.get('value',0)/1e6:.1f} M  qrf {q.get('value',0)/1e6:.1f} M  cpu {c.get('value',0)/1e6:.3f} M ({c.get('cores')} cores)  frac {r.get('frac')}  stale {r.get('ncu_stale')}  parity {j.get('parity_sample')} {q.get('parity_sample')}")
for k, v in (j.get("configs") or {}).items():
    print(k, json.dumps(v)[:300])
PY
bash tools/gpu_switch_matrix.sh
