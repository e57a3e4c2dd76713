#!/bin/bash
# bench.py --gpus N for several argument sets: tools/gpu_multi2.sh <tag> <N> "args1" "args2" ...
mkdir -p gpurun_out
TAG=$1; N=$2; shift; shift
q=0
for A in "$@"; do
  q=$((q+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29617+q)) bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline $A > gpurun_out/${TAG}_n${N}_$q.json 2> gpurun_out/${TAG}_n${N}_$q.err
  python - gpurun_out/${TAG}_n${N}_$q.json "$A" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("[%s] N=%d mesh %s ms/step %.2f value %.3e gs %.2f ms lu %.2f ms parity %s" % (sys.argv[2], d["n_gpus"], d["config"]["mesh"], d["ms_per_step"], d["value"], d["roofline"]["avg_launch_ms"], d["roofline"]["lu_avg_solve_ms"] or 0, (d.get("parity_check") or "")[:9]))
except Exception as e:
    print("FAILED", sys.argv[2], e, open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
