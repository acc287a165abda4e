#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden_pipeline.py tests/test_gpu_api.py tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -12
: > gpurun_out/r2q_sweep.jsonl
run() { env "$@" python tools/ktime.py 8192 20 60 2>> gpurun_out/r2q_err.log | tee -a gpurun_out/r2q_sweep.jsonl | cut -c1-330; }
run MJB_X=0
run MJB_LIB=build_ab/libmjb200_single.so
run MJB_X=0
run MJB_LIB=build_ab/libmjb200_single.so
timeout 300 python tools/diag_scene.py tendons 0 2>&1 | tail -30
