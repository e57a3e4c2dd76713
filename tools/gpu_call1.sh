#!/bin/bash
# first GPU pass of the round: parity tests, then the bench for the default build and the compile-time variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c1_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/c1_tests.log 2>&1
echo "tests rc $?" >> gpurun_out/c1_tests.log
B="--steps 3 --warmup 3 --no-cpu-baseline --no-as-configured"
timeout 300 python bench.py $B > gpurun_out/c1_bench_default.json 2> gpurun_out/c1_bench_default.err
for v in split2 pf3 d1 ty7pf3 pair4 b3ty12; do
  HYDRO_GPU_LIB=variants/libhydro_$v.so timeout 300 python bench.py $B > gpurun_out/c1_bench_$v.json 2> gpurun_out/c1_bench_$v.err
done
HYDRO_GT_CLOCK=1 HYDRO_GPU_LIB=variants/libhydro_clock.so timeout 300 python bench.py $B > gpurun_out/c1_bench_clock.json 2> gpurun_out/c1_bench_clock.err
tail -3 gpurun_out/c1_tests.log
for f in gpurun_out/c1_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.2f gs %.2f ms lu %.2f ms frac %.3f" % (d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["lu_avg_solve_ms"] or 0, d["roofline"]["whole_step"]["frac"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
cat gpurun_out/c1_bench_clock.err | tail -3
