#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_golden_pipeline.py -m gpu -q -k "implicit" 2>&1 | tail -40
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_golden_pipeline.py -m gpu -q -k "_implicit" > gpurun_out/r2y_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2y_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_golden_pipeline.py -m gpu -q -k "_implicit or mixed or humanoid" > gpurun_out/r2y_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -c "Race reported" gpurun_out/r2y_racecheck.log; tail -4 gpurun_out/r2y_racecheck.log
MJB_E2E_TRACE=1 timeout 300 python bench.py --no-cpu > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; cut -c1-200 gpurun_out/r2y_bench.json; grep e2e gpurun_out/r2y_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2y_bench.json')); print(d['value'], d['e2e'], d['kernel_ms'])"
