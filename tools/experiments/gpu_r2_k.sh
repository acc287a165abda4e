#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
: > gpurun_out/r2k_sweep.jsonl
for cls in 1 0; do for warm in 20 300; do
  MJB_SOLVER_CLASSES=$cls timeout 120 python tools/ktime.py 8192 $warm 100 >> gpurun_out/r2k_sweep.jsonl 2>>gpurun_out/r2k_err.log || echo "fail"
done; done
python - <<'PY'
import json
for l in open("gpurun_out/r2k_sweep.jsonl"):
  d=json.loads(l); c=d["cfg"]; print(c.get("MJB_SOLVER_CLASSES"), d["step_us"], d["Msteps_s"], d["kernel_us"], round(d["nefc"],1), round(d["niter"],2), d["nan"])
PY
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_suite.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/gpu_suite.log | cut -c1-240
