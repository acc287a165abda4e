#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
export MJB_SPLIT=1
K=${K:-k_solver}
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$K" -s ${SKIP:-20} -c ${CNT:-1} -o gpurun_out/r2e_prof -f python tools/prof_step.py 25 8192 > gpurun_out/r2e_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2e_ncu.log
ncu -i gpurun_out/r2e_prof.ncu-rep --page raw --csv > gpurun_out/r2e_raw.csv 2>/dev/null
ncu -i gpurun_out/r2e_prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2e_src.csv 2>/dev/null
