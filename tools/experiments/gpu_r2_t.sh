#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden_pipeline.py tests/test_gpu_api.py tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -6
timeout 2400 compute-sanitizer --tool racecheck --print-limit 2000 --error-exitcode 1 python -m pytest tests/test_gpu_golden_pipeline.py tests/test_gpu_api.py -q -m gpu -k "humanoid or mixed or actuators or tendons or mesh or equality or sensors or g1 or convex or stateful or batched or sparse or three" > gpurun_out/r2t_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -c "Race reported" gpurun_out/r2t_racecheck.log; grep "Race reported" gpurun_out/r2t_racecheck.log | sed 's/+0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -20; tail -3 gpurun_out/r2t_racecheck.log
python tools/ktime.py 8192 20 60 2>/dev/null | cut -c1-300
