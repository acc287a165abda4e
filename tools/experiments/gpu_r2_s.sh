#!/bin/bash
# full GPU suite + compute-sanitizer (memcheck, racecheck) over a subset of the golden / API tests
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8
K="humanoid or mixed or actuators or tendons or mesh or boxccd or equality or sensors or stateful or batched or sparse"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_golden_pipeline.py tests/test_gpu_api.py -q -m gpu -x -k "$K" > gpurun_out/r2s_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -c "Invalid\|misaligned" gpurun_out/r2s_memcheck.log; tail -4 gpurun_out/r2s_memcheck.log
timeout 1800 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_golden_pipeline.py -q -m gpu -x -k "humanoid or mixed or actuators or tendons or mesh or equality" > gpurun_out/r2s_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -c "hazard" gpurun_out/r2s_racecheck.log; tail -4 gpurun_out/r2s_racecheck.log
