#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2f_tests.log
: > gpurun_out/r2f_sweep.jsonl
export MJB_LPW_POS=8 MJB_LPW_VEL=8 MJB_WPB_POS=2 MJB_WPB_VEL=2
for warm in 20 300; do
timeout 120 python tools/ktime.py 8192 $warm 100 >> gpurun_out/r2f_sweep.jsonl 2>>gpurun_out/r2f_err.log || echo "fail"
MJB_LIB=build_ab/libmjb200_r01.so timeout 120 python tools/ktime.py 8192 $warm 100 >> gpurun_out/r2f_sweep.jsonl 2>>gpurun_out/r2f_err.log || echo "fail r01"
done
python - <<'PY'
import json
for l in open("gpurun_out/r2f_sweep.jsonl"):
  d=json.loads(l); c=d["cfg"]; print(c.get("MJB_LIB","new")[-12:], d["step_us"], d["Msteps_s"], d["kernel_us"], round(d["nefc"],1), round(d["niter"],2), d["nan"])
PY
