#!/bin/bash
# solver micro-variants one by one (same box), condim-3 fast path of k_constraint, fully implicit integrator + suites
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for warm in 20 200; do
  for v in solbase DEFAULT MULM SUM8 TCOPY1; do
    if [ $v = DEFAULT ]; then L=mujoco_warp_b200/libmjb200.so; else L=build_ab/libmjb200_$v.so; fi
    echo "== warm $warm $v"; MJB_LIB=$L python tools/ktime.py 8192 $warm 60 2>/dev/null | cut -c60-330
  done
done
timeout 1500 python -m pytest tests/test_gpu_golden_pipeline.py tests/test_gpu_parity.py tests/test_gpu_bench_config.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -15
