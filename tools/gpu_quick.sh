#!/bin/bash
# quick A/B: bench with environment variants given as arguments "tag:ENV=VAL ENV2=VAL2"
mkdir -p gpurun_out
B="--steps 3 --warmup 3 --no-cpu-baseline --no-as-configured"
for v in "$@"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python bench.py $B > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err
  python - gpurun_out/q_$tag.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.2f gs %.2f ms lu %.2f ms frac %.3f e2e %.1f" % (d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["lu_avg_solve_ms"] or 0, d["roofline"]["whole_step"]["frac"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1].replace(".json",".err")).read()[-400:])
PY
done
