#!/bin/bash
# build_variant.sh NAME [-DMACRO=VALUE ...]  ->  build/variants/libcarenv_NAME.so  (kernel tuning experiments;
# run with CARENV_LIB=build/variants/libcarenv_NAME.so python benchmarks/ab_step.py ...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --shared -Xcompiler -fPIC \
  -Xcompiler -ffp-contract=off -std=c++17 "$@" -o build/variants/libcarenv_$name.so ppo_car_b200/csrc/carenv_kernels.cu
echo built build/variants/libcarenv_$name.so
