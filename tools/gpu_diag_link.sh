#!/bin/bash
# one GPU: the slab instantiation of the sweep kernel without communication (HYDRO_GT_FORCE_LINK) against the plain one, with the
# cycle counters of the instrumented build
mkdir -p gpurun_out
B="--steps 3 --warmup 3 --no-cpu-baseline --no-as-configured"
run() { tag=$1; shift; env "$@" timeout 300 python bench.py $B > gpurun_out/dl_$tag.json 2> gpurun_out/dl_$tag.err
  python - gpurun_out/dl_$tag.json $tag <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-14s ms/step %.2f gs %.2f ms lu %.2f ms" % (sys.argv[2], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["lu_avg_solve_ms"] or 0))
except Exception as e:
    print(sys.argv[2], "FAILED", e, open(sys.argv[1].replace(".json",".err")).read()[-300:])
PY
}
run plain X=1
run flink HYDRO_GT_FORCE_LINK=1
run clk HYDRO_GT_CLOCK=1 HYDRO_GPU_LIB=variants/libhydro_clock.so
run clk_flink HYDRO_GT_CLOCK=1 HYDRO_GT_FORCE_LINK=1 HYDRO_GPU_LIB=variants/libhydro_clock.so
grep -h gt_clocks gpurun_out/dl_*.err
