#!/bin/bash
# bench of the default build and of the variants named on the command line (variants/libhydro_<name>.so)
mkdir -p gpurun_out
TAG=$1; shift
B="--steps 3 --warmup 3 --no-cpu-baseline --no-as-configured"
timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
for v in "$@"; do
  HYDRO_GT_CLOCK=1 HYDRO_GPU_LIB=variants/libhydro_$v.so timeout 300 python bench.py $B > gpurun_out/${TAG}_bench_$v.json 2> gpurun_out/${TAG}_bench_$v.err
done
for f in gpurun_out/${TAG}_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.2f gs %.2f ms lu %.2f ms frac %.3f e2e %.1f" % (d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["lu_avg_solve_ms"] or 0, d["roofline"]["whole_step"]["frac"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1].replace(".json",".err")).read()[-300:])
PY
done
grep -h gt_clocks gpurun_out/${TAG}_bench_*.err
