#!/bin/bash
# Builds compile-time variants of the CUDA library into variants/ (git-ignored; travels to the GPU box) for A/B timing:
#   tools/build_variants.sh "name:-DFLAG=.. -DFLAG2=.." ...      then  HYDRO_GPU_LIB=variants/libhydro_<name>.so python bench.py
cd "$(dirname "$0")/../hydro_b200/csrc" || exit 1
mkdir -p ../../variants
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -shared -Xcompiler -fPIC $flags \
       -o ../../variants/libhydro_$name.so hydro_gpu.cu > ../../variants/build_$name.log 2>&1 &
done
wait
grep -l "error" ../../variants/build_*.log
exit 0
