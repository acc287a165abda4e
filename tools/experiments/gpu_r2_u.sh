#!/bin/bash
# HEAD validation: full GPU suite, smoke, humanoid bench line, kernel timing at the bench state
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2u_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2u_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/r2u_bench.json
python tools/ktime.py 8192 20 60 2>/dev/null | cut -c1-400
