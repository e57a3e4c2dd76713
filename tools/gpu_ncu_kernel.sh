#!/bin/bash
# one `ncu --set full` capture of a kernel inside the bench workload:  tools/gpu_ncu_kernel.sh <tag> <kernel regex> [skip] [lib]
mkdir -p gpurun_out
TAG=$1; KRE=$2; SKIP=${3:-3}; LIB=${4:-}
[ -n "$LIB" ] && export HYDRO_GPU_LIB=$LIB
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -f -o gpurun_out/$TAG \
   python bench.py --size 256 --steps 1 --warmup 3 --no-cpu-baseline --no-as-configured > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/$TAG.ncu-rep
