#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r2g_tests.log
