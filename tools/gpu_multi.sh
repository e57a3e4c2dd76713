#!/bin/bash
# multi-GPU pass: slab tests on separate devices, then bench.py --gpus N (weak; optional extra args): tools/gpu_multi.sh <tag> <N> [bench args]
mkdir -p gpurun_out
TAG=$1; N=$2; shift; shift
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc $?" >> gpurun_out/${TAG}_tests.log
tail -6 gpurun_out/${TAG}_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus $N --steps 3 --warmup 3 "$@" > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
tail -c 400 gpurun_out/${TAG}_bench_n$N.err
python - gpurun_out/${TAG}_bench_n$N.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("N=%d ms/step %.2f value %.3e gs %.2f ms lu %.2f ms parity %s kernel %s" % (d["n_gpus"], d["ms_per_step"], d["value"], d["roofline"]["avg_launch_ms"], d["roofline"]["lu_avg_solve_ms"] or 0, d.get("parity_check"), d["roofline"]["kernel"][:30]))
except Exception as e:
    print("FAILED", e)
PY
