#!/bin/bash
# parity tests + bench of the default build and of variants:  tools/gpu_test_bench.sh <tag> [variants...]
mkdir -p gpurun_out
TAG=$1; shift
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc $?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
bash tools/gpu_bench_variants.sh $TAG "$@"
