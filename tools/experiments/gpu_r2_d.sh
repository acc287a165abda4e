#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2d_tests.log
: > gpurun_out/r2d_sweep.jsonl
export MJB_LPW_POS=8 MJB_LPW_VEL=8 MJB_WPB_POS=2 MJB_WPB_VEL=2
MJB_LIB=build_ab/libmjb200_r01.so timeout 120 python tools/ktime.py 8192 20 100 >> gpurun_out/r2d_sweep.jsonl 2>>gpurun_out/r2d_err.log || echo "fail r01"
for cfg in "1 1 2" "2 2 2" "2 2 3" "2 2 4" "2 2 1"; do
  set -- $cfg
  MJB_WPB_COL=$1 MJB_WPB_CON=$2 MJB_SPLIT=$3 timeout 120 python tools/ktime.py 8192 20 100 >> gpurun_out/r2d_sweep.jsonl 2>>gpurun_out/r2d_err.log || echo "fail $cfg"
done
python - <<'PY'
import json
for l in open("gpurun_out/r2d_sweep.jsonl"):
  d=json.loads(l); c=d["cfg"]; print(c.get("MJB_LIB","new")[-12:], c.get("MJB_WPB_COL"), c.get("MJB_WPB_CON"), c.get("MJB_SPLIT"), d["step_us"], d["kernel_us"], d["nefc"], d["niter"], d["nan"])
PY
