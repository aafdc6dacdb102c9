#!/usr/bin/env bash
# r02_w: the whole C5 grid's worth of work on one B200, device-resident: 92 steps of 5 x 10^7 pixels x 8 days
#        = 3.68 x 10^10 pixel-days = 10^8 pixels x 368 days (BASELINE.json configs[4]); the band of 5000 raster rows moves down
#        the 10^4 x 10^4 raster from step to step, every step pays its own geometry pass.
#   gpurun --timeout 1200 -- 'bash tools/gpu_r02w.sh'
set -u
mkdir -p gpurun_out
FC=/tmp/rfsi_bench_forest.npz
timeout 900 python bench.py --rows 5000 --days 8 --steps 92 --warmup 3 --no-cpu --no-e2e --no-configs --no-qrf --no-inprocess --forest-cache $FC > gpurun_out/w_fullgrid.json 2> gpurun_out/w_fullgrid.err; echo "full grid rc=$?"
tail -n 3 gpurun_out/w_fullgrid.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/w_fullgrid.json").read().strip().splitlines()[-1]); r = j["roofline"]
print(f"value {j['value']/1e6:.1f} M pixel-days/s  steps {j['steps']} x {j['ms_per_step']:.1f} ms = {j['steps']*j['ms_per_step']/1e3:.1f} s for {j['steps']*j['config']['pixels_per_step']*j['config']['days_per_step']:.3e} pixel-days  clocks {j.get('clocks')}")
PY
