#!/bin/bash
# A/B of library variants on N GPUs (and on one): tools/gpu_ab_lib.sh <N> name1 name2 ...   (name "default" = the in-tree library)
mkdir -p gpurun_out
N=$1; shift
q=0
for v in "$@"; do
  q=$((q+1))
  LIB=variants/libhydro_$v.so; [ "$v" = default ] && LIB=hydro_b200/libhydro_gpu.so
  if [ "$N" = 1 ]; then
    HYDRO_GPU_LIB=$LIB timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-as-configured > gpurun_out/ab_${v}_n1.json 2> gpurun_out/ab_${v}_n1.err
  else
    HYDRO_GPU_LIB=$LIB timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29717+q)) bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-as-configured > gpurun_out/ab_${v}_n$N.json 2> gpurun_out/ab_${v}_n$N.err
  fi
  python - gpurun_out/ab_${v}_n$N.json $v <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("%-12s N=%d ms/step %.2f value %.3e gs %.2f ms lu %.2f ms parity %s" % (sys.argv[2], d["n_gpus"], d["ms_per_step"], d["value"], d["roofline"]["avg_launch_ms"], d["roofline"]["lu_avg_solve_ms"] or 0, (d.get("parity_check") or "")[:9]))
except Exception as e:
    print(sys.argv[2], "FAILED", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
