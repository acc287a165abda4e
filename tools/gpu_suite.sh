#!/bin/bash
# full GPU test suite + a short humanoid kernel timing (tools/ktime.py)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_suite.log 2>&1; echo "tests rc=$?"; tail -${TAIL:-12} gpurun_out/gpu_suite.log | cut -c1-240
timeout 120 python tools/ktime.py 8192 20 100 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['step_us'], d['Msteps_s'], d['kernel_us'])"
