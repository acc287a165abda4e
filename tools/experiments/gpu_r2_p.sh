#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
: > gpurun_out/r2p_sweep.jsonl
run() { env "$@" python tools/ktime.py 8192 20 60 2>> gpurun_out/r2p_err.log | tee -a gpurun_out/r2p_sweep.jsonl | cut -c1-330; }
run MJB_X=0
run MJB_LPW_POS=16 MJB_WPB_POS=4
run MJB_LPW_POS=16 MJB_WPB_POS=2
run MJB_LPW_VEL=16 MJB_WPB_VEL=4
run MJB_LPW_VEL=16 MJB_WPB_VEL=2
run MJB_LPW_POS=8 MJB_WPB_POS=4 MJB_LPW_VEL=8 MJB_WPB_VEL=4
run MJB_LPW_POS=4 MJB_WPB_POS=2 MJB_LPW_VEL=4 MJB_WPB_VEL=2
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_golden_pipeline.py tests/test_gpu_parity.py tests/test_gpu_bench_config.py -q -m gpu 2>&1 | tail -8
