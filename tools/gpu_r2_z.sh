#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
for warm in 20 200; do
  echo "== pre (plain, unrolled stager fallbacks) $warm"; MJB_LIB=build_ab/libmjb200_pre.so python tools/ktime.py 8192 $warm 60 2>/dev/null | cut -c1-330
  echo "== default (nreg, rolled stager fallbacks) $warm"; python tools/ktime.py 8192 $warm 60 2>/dev/null | cut -c1-330
  echo "== constpad $warm"; MJB_LIB=build_ab/libmjb200_constpad.so python tools/ktime.py 8192 $warm 60 2>/dev/null | cut -c1-330
done
MJB_LIB=build_ab/libmjb200_constpad.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
