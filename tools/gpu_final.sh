#!/bin/bash
# evidence pass of the round on one GPU: default bench (with CPU baseline and the as-configured leg), reference arm, launch list,
# one ncu --set full capture of the sweep kernel and of the lu kernel: tools/gpu_final.sh <tag>
mkdir -p gpurun_out
TAG=$1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 200 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_reference.json 2> gpurun_out/${TAG}_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-as-configured > gpurun_out/${TAG}_ncu_bench.log 2>&1
bash tools/gpu_ncu_multi.sh ${TAG}_gs "k_gs_tiled" 3 1
bash tools/gpu_ncu_multi.sh ${TAG}_lu "k_lu_tiled" 6 1
python - gpurun_out/${TAG}_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.2f gs %.2f ms lu %.2f ms frac %.3f e2e %.1f cpu %s asconf %s" % (d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["lu_avg_solve_ms"] or 0, d["roofline"]["whole_step"]["frac"], d["e2e"]["ms_per_step"], (d.get("cpu_baseline") or {}).get("value"), (d.get("as_configured") or {}).get("ms_per_step")))
PY
bash tools/gpu_ncu_multi.sh ${TAG}_fast "k_f[a-e]_" 12 6
