#!/bin/bash
# parity tests, default bench, launch list (ncu gpu__time_duration) of one bench run
mkdir -p gpurun_out
TAG=$1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc $?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-as-configured > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-as-configured > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bench.log | cut -c1-300
