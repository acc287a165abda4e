#!/bin/bash
# ncu --set full on the team kernels (position, velocity) at LPW=8 WPB=2, with source-level stall attribution
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
export MJB_LPW_POS=${LPW:-8} MJB_LPW_VEL=${LPW:-8} MJB_WPB_POS=${WPB:-2} MJB_WPB_VEL=${WPB:-2} MJB_SPLIT=1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'k_position|k_velocity' -s 40 -c 2 -o gpurun_out/r2c_prof -f python tools/prof_step.py 25 8192 > gpurun_out/r2c_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2c_ncu.log
ncu -i gpurun_out/r2c_prof.ncu-rep --page raw --csv > gpurun_out/r2c_raw.csv 2>/dev/null
ncu -i gpurun_out/r2c_prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2c_src.csv 2>/dev/null
ls -la gpurun_out/r2c_*
