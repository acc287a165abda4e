#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden_pipeline.py tests/test_gpu_parity.py tests/test_gpu_bench_config.py tests/test_gpu_api.py -q -m gpu > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2m_tests.log
python tools/ktime.py 8192 20 100 2> gpurun_out/r2m_err.log | tee gpurun_out/r2m_sweep.jsonl
python tools/ktime.py 8192 20 100 2>> gpurun_out/r2m_err.log | tee -a gpurun_out/r2m_sweep.jsonl
timeout 600 python bench.py --no-cpu > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"; tail -c 1800 gpurun_out/r2m_bench.json
