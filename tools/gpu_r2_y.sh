#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_golden_pipeline.py -m gpu -q -x -k "tendons_implicit" 2>&1 | tail -40
timeout 900 python -m pytest tests/test_gpu_api.py -m gpu -q 2>&1 | tail -8
python tools/ktime.py 8192 200 60 2>/dev/null | cut -c1-330
