#!/usr/bin/env bash
# First GPU call of a round, meant to run under gpurun on ONE B200 (about 12 minutes of box time):
#   gpurun --timeout 1200 -- 'bash tools/gpu_round_start.sh'
# 1. the whole GPU suite + smoke + the default bench (must all exit 0 before anything is profiled);
# 2. A/B of the forest images and of J / P on the same box (same forest, cached between runs);
# 3. the launch list of the bench (share of each kernel in a step);
# 4. one `--set full` capture of the fused kernel and of the quantile kernel.
# Everything lands in gpurun_out/ (scratch); copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
FC=/tmp/rfsi_bench_forest.npz     # 1.8 GB: NOT under gpurun_out/ (64 MiB come back at most)
B="python bench.py --no-cpu --qrf --forest-cache $FC"
if [ "${SKIP_SUITE:-0}" = "1" ]; then
  python -m pytest tests/test_gpu_r_shim.py tests/test_gpu_fork.py tests/test_gpu_forest.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
else
  python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
fi
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
python tools/host_copy_bench.py > gpurun_out/host_copy_bench.txt 2>&1
RFSI_B200_TRACE=1 python tests/tools/config_timings.py > gpurun_out/config_timings.md 2> gpurun_out/config_timings.err
$B > gpurun_out/bench_sector.json 2> gpurun_out/bench_sector.err; echo "bench rc=$?" | tee -a gpurun_out/bench_sector.err
RFSI_B200_COMPACT=2 $B > gpurun_out/bench_blocked.json 2> gpurun_out/bench_blocked.err; echo "bench rc=$?" | tee -a gpurun_out/bench_blocked.err
for v in "J=12" "J=16" "J=4" "P=128" "P=128 J=16" "P=128 J=12"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $(for kv in $v; do echo "RFSI_B200_$kv"; done) $B --no-e2e --steps 4 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  echo "variant $v rc=$?" | tee -a gpurun_out/bench_$tag.err
done
grep -q "rc=0" gpurun_out/pytest_gpu.log && grep -q "rc=0" gpurun_out/bench_sector.err || { echo "not profiling a failing build"; ls -la gpurun_out; exit 1; }
P="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --qrf --forest-cache $FC"
$P > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $P > gpurun_out/ncu_launches.log 2>&1
P1="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --qrf --forest-cache $FC"
$P1 > gpurun_out/plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 1 -o gpurun_out/fused_sector -f $P1 > gpurun_out/ncu_fused.log 2>&1
$P1 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qrf_quantile -c 1 -o gpurun_out/qrf_direct -f $P1 > gpurun_out/ncu_qrf.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        q = j.get("qrf") or {}
        e = j.get("e2e") or {}
        r = j["roofline"]
        print(f"{f[17:-5]:14s} value {j['value']/1e6:7.1f} M  e2e {e.get('value', 0)/1e6:7.1f} M  qrf {q.get('value', 0)/1e6:6.1f} M "
              f"(q-kernel {q.get('quantile_kernel_ms_per_step', 0):.2f} ms, fused {q.get('fused_kernel_ms_per_step', 0):.2f} ms)  "
              f"fused launch {r['avg_launch_ms']:.2f} ms  d {r['mean_internal_nodes_per_tree']:.2f}  forest {r['forest_bytes']/1e6:.0f} MB")
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
du -sh gpurun_out; ls -la gpurun_out
