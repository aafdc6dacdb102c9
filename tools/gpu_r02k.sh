#!/usr/bin/env bash
# r02_k: host chunk size of the fused pipeline (e2e) and days per quantile launch (scratch budget), one box.
#   gpurun --timeout 1200 -- 'bash tools/gpu_r02k.sh'
set -u
mkdir -p gpurun_out
FC=/tmp/rfsi_bench_forest.npz
B="python bench.py --no-cpu --no-configs --steps 4 --warmup 3 --forest-cache $FC"
$B --no-qrf > gpurun_out/k_chunk_1M.json 2> gpurun_out/k_chunk_1M.err; echo "1M rc=$?"
RFSI_B200_FUSED_CHUNK=524288 $B --no-qrf > gpurun_out/k_chunk_512k.json 2> gpurun_out/k_chunk_512k.err; echo "512k rc=$?"
RFSI_B200_FUSED_CHUNK=262144 $B --no-qrf > gpurun_out/k_chunk_256k.json 2> gpurun_out/k_chunk_256k.err; echo "256k rc=$?"
RFSI_B200_FUSED_CHUNK=2097152 $B --no-qrf > gpurun_out/k_chunk_2M.json 2> gpurun_out/k_chunk_2M.err; echo "2M rc=$?"
RFSI_B200_QRF_SCRATCH_MB=8192 $B --no-e2e > gpurun_out/k_scratch_8G.json 2> gpurun_out/k_scratch_8G.err; echo "8G rc=$?"
$B --no-e2e > gpurun_out/k_scratch_1G.json 2> gpurun_out/k_scratch_1G.err; echo "1G rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/k_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1]); q = j.get('qrf') or {}; e = j.get('e2e') or {}
        print(f"{f[11:-5]:14s} value {j['value']/1e6:7.1f} M  e2e {e.get('value', 0)/1e6:7.1f} M ({e.get('ms_per_step', 0):.2f} ms)  qrf {q.get('value', 0)/1e6:.1f} M "
              f"(fused {q.get('fused_kernel_ms_per_step', 0):.2f} + q-kernel {q.get('quantile_kernel_ms_per_step', 0):.2f} ms)")
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
