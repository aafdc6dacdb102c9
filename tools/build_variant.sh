#!/usr/bin/env bash
# Experiment builds of the engine: tools/build_variant.sh <name> [-DMACRO=1 ...] -> rfsi_b200/librfsi_b200_<name>.so
# (select it with RFSI_B200_LIB=rfsi_b200/librfsi_b200_<name>.so; built .so files travel with gpurun, not with git)
set -eu
name=$1; shift
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -shared -Xcompiler -fPIC -cudart static \
     "$@" -o rfsi_b200/librfsi_b200_$name.so rfsi_b200/csrc/engine.cu
