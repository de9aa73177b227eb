#!/bin/bash
# Development aid: build a differently-compiled copy of the library for A/B runs on the GPU box.
#   tools/build_variant.sh NAME [-DFLAG ...]   ->  gpurun_variants/libmfg_NAME.so   (use with MFG_B200_LIB=...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p gpurun_variants/obj_$name
for f in mfg_abi mfg_obs mfg_step; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC "$@" \
    -c marl_factory_grid_b200/csrc/$f.cu -o gpurun_variants/obj_$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a --shared -o gpurun_variants/libmfg_$name.so gpurun_variants/obj_$name/*.o
rm -rf gpurun_variants/obj_$name
echo built gpurun_variants/libmfg_$name.so
