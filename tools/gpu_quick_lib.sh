#!/bin/bash
# A/B of library variants with the same environment: tools/gpu_quick_lib.sh "ENV=.." name1 name2 ...  (name "default" = the in-tree library)
mkdir -p gpurun_out
ENVS=$1; shift
for v in "$@"; do
  LIB=variants/libhydro_$v.so; [ "$v" = default ] && LIB=hydro_b200/libhydro_gpu.so
  env $ENVS HYDRO_GPU_LIB=$LIB timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-as-configured > gpurun_out/ql_$v.json 2> gpurun_out/ql_$v.err
  python - gpurun_out/ql_$v.json $v <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-12s ms/step %.2f gs %.2f ms lu %.2f ms" % (sys.argv[2], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["lu_avg_solve_ms"] or 0))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
