#!/bin/bash
for v in "$@"; do
  cp variants/lib$v.so hydro_b200/libhydro_gpu.so
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; print('$v', 'step %.2f ms' % d['ms_per_step'], 'gs %.2f ms' % r['avg_launch_ms'], 'lu %.2f ms' % (r['lu_kernel_share_of_step']*d['ms_per_step']/3))
except Exception as e: print('$v', 'FAILED', e)
"
done
