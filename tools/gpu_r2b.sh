#!/bin/bash
# targeted tests (-k expression), default bench, launch list
mkdir -p gpurun_out
TAG=$1; KEXPR=$2
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x -k "$KEXPR" > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc $?" >> gpurun_out/${TAG}_tests.log
tail -15 gpurun_out/${TAG}_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-as-configured > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 300 gpurun_out/${TAG}_bench.err
python - gpurun_out/${TAG}_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.2f gs %.2f ms lu %.2f ms frac %.3f e2e %.1f" % (d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["lu_avg_solve_ms"] or 0, d["roofline"]["whole_step"]["frac"], d["e2e"]["ms_per_step"]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-as-configured > gpurun_out/${TAG}_ncu_bench.log 2>&1
