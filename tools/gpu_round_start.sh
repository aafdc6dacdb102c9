#!/usr/bin/env bash
# First GPU call of a round, meant to run under gpurun on ONE B200 (about 6 minutes of box time):
#   gpurun --timeout 600 -- 'bash tools/gpu_round_start.sh'
# 1. the whole GPU suite + smoke + the default bench (must all exit 0 before anything is profiled);
# 2. the launch list of the bench (share of each kernel in a step);
# 3. one `--set full` capture of each quantile kernel form: the rank form that is
#    the default and the strip form that measured slower -- read them with tools/ncu_summary.py / tools/ncu_lines.py.
# Everything lands in gpurun_out/ (scratch); copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
python bench.py --qrf > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" | tee -a gpurun_out/bench_n1.err
grep -q "rc=0" gpurun_out/pytest_gpu.log && grep -q "rc=0" gpurun_out/bench_n1.err || { echo "not profiling a failing build"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --qrf > gpurun_out/ncu_launches.log 2>&1
for strip in 0 1; do
  RFSI_B200_QRF_STRIP=$strip ncu --set full --clock-control none --import-source on -k regex:qrf_quantile -c 1 \
      -o gpurun_out/qrf_strip$strip -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --qrf \
      > gpurun_out/ncu_qrf_strip$strip.log 2>&1
done
ls -la gpurun_out
