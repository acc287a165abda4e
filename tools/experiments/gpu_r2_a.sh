#!/bin/bash
# round-2 experiment A: team kernels (position / velocity) -- parity, then LPW sweep, then the full GPU suite
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2a_parity.log 2>&1; echo "parity rc=$?"
tail -5 gpurun_out/r2a_parity.log
for lpw in 8 4 16 32; do
  MJB_LPW_POS=$lpw MJB_LPW_VEL=$lpw timeout 300 python bench.py --steps 100 --warmup 20 --no-cpu > gpurun_out/r2a_bench_$lpw.json 2> gpurun_out/r2a_bench_$lpw.err
  echo "lpw=$lpw rc=$?"; python - <<PY
import json
try:
  d=json.loads(open("gpurun_out/r2a_bench_$lpw.json").read().strip().splitlines()[-1])
  print($lpw, d["value"], d["ms_per_step"], d["kernel_ms"], d["config"]["ncon_mean"], d["config"]["nefc_mean"], d["config"]["solver_niter_mean"])
except Exception as e: print("fail", e)
PY
done
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_gpu_tests.log 2>&1; echo "suite rc=$?"
tail -15 gpurun_out/r2a_gpu_tests.log
