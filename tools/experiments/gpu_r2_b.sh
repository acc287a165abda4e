#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -m gpu > gpurun_out/r2b_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r2b_parity.log
: > gpurun_out/r2b_sweep.jsonl
for cfg in ${SWEEP:-"8 1" "8 2" "8 4" "16 2" "16 4" "32 2" "4 2"}; do
  set -- $cfg
  MJB_LPW_POS=$1 MJB_LPW_VEL=$1 MJB_WPB_POS=$2 MJB_WPB_VEL=$2 timeout 120 python tools/ktime.py 8192 20 100 >> gpurun_out/r2b_sweep.jsonl 2>gpurun_out/r2b_err.log || echo "fail $cfg"
done
python - <<'PY'
import json
for l in open("gpurun_out/r2b_sweep.jsonl"):
  d=json.loads(l); print(d["cfg"].get("MJB_LPW_POS"), d["cfg"].get("MJB_WPB_POS"), d["step_us"], d["kernel_us"], d["nan"])
PY
