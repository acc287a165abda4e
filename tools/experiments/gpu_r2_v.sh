#!/bin/bash
# A/B on one box: solver baseline vs mul_m_small + warp_sum8, velocity fork, split count; per-line instruction profile of the solver
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
B=build_ab/libmjb200_solbase.so
echo "== base solver, serial chain";   MJB_LIB=$B MJB_FORK=0 python tools/ktime.py 8192 20 60 2>/dev/null | cut -c1-330
echo "== new solver, serial chain";    MJB_FORK=0 python tools/ktime.py 8192 20 60 2>/dev/null | cut -c1-330
echo "== new solver, velocity fork";   MJB_FORK=1 python tools/ktime.py 8192 20 60 2>/dev/null | cut -c1-330
for sp in 1 3 4; do echo "== fork, split $sp"; MJB_FORK=1 MJB_SPLIT=$sp python tools/ktime.py 8192 20 60 2>/dev/null | cut -c1-120; done
echo "== base solver, fork"; MJB_LIB=$B MJB_FORK=1 python tools/ktime.py 8192 20 60 2>/dev/null | cut -c1-120
echo "== bench-state (200 steps in) new+fork vs base serial"
MJB_FORK=1 python tools/ktime.py 8192 200 60 2>/dev/null | cut -c1-330
MJB_LIB=$B MJB_FORK=0 python tools/ktime.py 8192 200 60 2>/dev/null | cut -c1-330
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_config.py tests/test_gpu_golden_pipeline.py -m gpu -q -x 2>&1 | tail -4
MJB_SPLIT=1 timeout 600 ncu --section SourceCounters --import-source on --clock-control none -k regex:k_solver -s 20 -c 1 -o gpurun_out/r2v_sol -f python tools/prof_step.py 25 8192 > gpurun_out/r2v_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2v_sol.ncu-rep --page source --csv --print-source cuda > gpurun_out/r2v_sol_src.csv 2>/dev/null
ncu -i gpurun_out/r2v_sol.ncu-rep --page raw --csv > gpurun_out/r2v_sol_raw.csv 2>/dev/null
ls -la gpurun_out/r2v_*
