#!/bin/bash
# bench of the default build, then one ncu --set full capture of a kernel:  tools/gpu_bench_ncu.sh <tag> <kernel regex> [skip]
bash tools/gpu_bench_variants.sh $1
bash tools/gpu_ncu_kernel.sh ${1}_ncu $2 ${3:-6}
