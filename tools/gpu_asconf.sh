#!/bin/bash
# the as-configured leg of bench.py with the sweep-chunk trace: tools/gpu_asconf.sh "tag:ENV=VAL ..." ...
mkdir -p gpurun_out
for v in "$@"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs HYDRO_SOR_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ac_$tag.json 2> gpurun_out/ac_$tag.err
  python - gpurun_out/ac_$tag.json $tag <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])["as_configured"]
print(sys.argv[2], {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d.items() if k != "workload"})
PY
done
