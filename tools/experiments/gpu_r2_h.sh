#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r2h_tests.log | cut -c1-250
for wl in convex_mesh g1 three_humanoids; do
  timeout 300 python bench.py --workload $wl --steps 50 --warmup 20 --no-cpu > gpurun_out/r2h_bench_$wl.json 2> gpurun_out/r2h_bench_$wl.err; echo "$wl rc=$?"; tail -3 gpurun_out/r2h_bench_$wl.err
  python - <<PY
import json
try:
  d=json.loads(open("gpurun_out/r2h_bench_$wl.json").read().strip().splitlines()[-1])
  print("$wl", round(d["value"]/1e6,3), "M steps/s", d["ms_per_step"], d["kernel_ms"], {k:d["config"][k] for k in ("ncon_mean","nefc_mean","solver_niter_mean","overflow_worlds","nan_worlds")}, d["e2e"]["value"])
except Exception as e: print("fail", e)
PY
done
