#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
: > gpurun_out/r2j_sweep.jsonl
for big in 0 1; do for jc in 0 32; do
  [ "$big" = 0 ] && [ "$jc" = 32 ] && continue
  MJB_SOLVER_BIG=$big MJB_JCAP=$jc timeout 120 python tools/ktime.py 8192 20 100 >> gpurun_out/r2j_sweep.jsonl 2>>gpurun_out/r2j_err.log || echo "fail"
  MJB_SOLVER_BIG=$big MJB_JCAP=$jc timeout 120 python tools/ktime.py 8192 300 100 >> gpurun_out/r2j_sweep.jsonl 2>>gpurun_out/r2j_err.log || echo "fail"
done; done
python - <<'PY'
import json
for l in open("gpurun_out/r2j_sweep.jsonl"):
  d=json.loads(l); c=d["cfg"]; print(c.get("MJB_SOLVER_BIG"), c.get("MJB_JCAP"), d["step_us"], d["Msteps_s"], d["kernel_us"], round(d["nefc"],1), round(d["niter"],2), d["nan"])
PY
MJB_SOLVER_BIG=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden_pipeline.py -x -q -m gpu 2>&1 | tail -3
