#!/bin/bash
# bench.py --gpus N with environment variants: tools/gpu_n2_env.sh <N> "tag:ENV=VAL ENV2=VAL2" ...
mkdir -p gpurun_out
N=$1; shift; q=0
for v in "$@"; do
  q=$((q+1)); tag=${v%%:*}; envs=${v#*:}
  env $envs HYDRO_BENCH_RANKS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29817+q)) bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-as-configured > gpurun_out/ne_${tag}_n$N.json 2> gpurun_out/ne_${tag}_n$N.err
  python - gpurun_out/ne_${tag}_n$N.json $tag <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("%-12s N=%d ms/step %.2f value %.3e gs %.2f ms lu %.2f ms parity %s" % (sys.argv[2], d["n_gpus"], d["ms_per_step"], d["value"], d["roofline"]["avg_launch_ms"], d["roofline"]["lu_avg_solve_ms"] or 0, (d.get("parity_check") or "")[:9]))
except Exception as e:
    print(sys.argv[2], "FAILED", e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
  grep -h "^rank" gpurun_out/ne_${tag}_n$N.err
done
