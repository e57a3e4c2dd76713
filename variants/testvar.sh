#!/bin/bash
for v in "$@"; do
  cp variants/lib$v.so hydro_b200/libhydro_gpu.so
  echo "== $v: $(timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -1)"
done
