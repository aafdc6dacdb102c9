#!/usr/bin/env bash
# The GPU parity tests once per engine switch (every non-default code path stays bit-equal to the oracle).
#   gpurun --timeout 2400 -- 'bash tools/gpu_switch_matrix.sh'   -> gpurun_out/switch_matrix.txt
set -u
mkdir -p gpurun_out
OUT=gpurun_out/switch_matrix.txt
echo "python -m pytest tests/test_gpu_{forest,fused,fuzz,maps,knn,api,multi,raster}.py -m gpu -q on one B200, once per engine switch" > $OUT
T="tests/test_gpu_forest.py tests/test_gpu_fused.py tests/test_gpu_fuzz.py tests/test_gpu_maps.py tests/test_gpu_knn.py tests/test_gpu_api.py tests/test_gpu_multi.py tests/test_gpu_raster.py"
run() {
  local line
  line=$(env "$@" python -m pytest $T -m gpu -q 2>&1 | tail -n 1)
  printf "%-44s %s\n" "$*" "$line" | tee -a $OUT
}
run RFSI_B200_COMPACT=3
run RFSI_B200_COMPACT=2
run RFSI_B200_COMPACT=1
run RFSI_B200_COMPACT=0
run RFSI_B200_QRF_FAST=0
run RFSI_B200_QRF_RANKS=0 RFSI_B200_COMPACT=2
run RFSI_B200_QRF_SCRATCH_MB=64
run RFSI_B200_FUSED_CHUNK=4096
run RFSI_B200_HOST_THREADS=1
run RFSI_B200_KNN_F32_PREFILTER=0
run RFSI_B200_TILE2D=0
run RFSI_B200_MORTON=0
run RFSI_B200_CELLS_MIN=1
run RFSI_B200_CELLS_MIN=1000000
run RFSI_B200_J=4
run RFSI_B200_J=12
run RFSI_B200_P=128
run RFSI_B200_KC_EXTRA=0
run RFSI_B200_CHUNK_MB=1
run RFSI_B200_RANK_CACHE=0
run RFSI_B200_RANK_BUCKETS=64
run RFSI_B200_RANK_BUCKETS=4096 RFSI_B200_RANK_CACHE=0
