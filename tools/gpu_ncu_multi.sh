#!/bin/bash
# ncu --set full of several launches matching a regex inside the bench workload: tools/gpu_ncu_multi.sh <tag> <regex> <skip> <count>
mkdir -p gpurun_out
TAG=$1; KRE=$2; SKIP=${3:-0}; CNT=${4:-4}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c $CNT -f -o gpurun_out/$TAG \
   python bench.py --size 256 --steps 1 --warmup 3 --no-cpu-baseline --no-as-configured > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log | cut -c1-200
ls -la gpurun_out/$TAG.ncu-rep
