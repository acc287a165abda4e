#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
for warm in 20 200; do
  echo "== default $warm"; python tools/ktime.py 8192 $warm 60 2>/dev/null | cut -c1-330
  echo "== plain $warm"; MJB_LIB=build_ab/libmjb200_plain.so python tools/ktime.py 8192 $warm 60 2>/dev/null | cut -c1-330
done
MJB_LIB=build_ab/libmjb200_plain.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_config.py -m gpu -q -x 2>&1 | tail -3
