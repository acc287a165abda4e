#!/bin/bash
# A/B on one box: r02 library (solver + constraint of HEAD~) vs the register-capped solver with mul_m_small / warp_sum8 and the
# row-phase k_constraint; parity suites; per-line profile of the non-solver kernels at the contact-rich bench state
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for warm in 20 200; do
  echo "== warm $warm: base";            MJB_LIB=build_ab/libmjb200_base.so python tools/ktime.py 8192 $warm 60 2>/dev/null | cut -c1-330
  echo "== warm $warm: new solver only"; MJB_LIB=build_ab/libmjb200_conbase.so python tools/ktime.py 8192 $warm 60 2>/dev/null | cut -c1-330
  echo "== warm $warm: new";             python tools/ktime.py 8192 $warm 60 2>/dev/null | cut -c1-330
done
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_config.py tests/test_gpu_golden_pipeline.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -6
timeout 300 python bench.py --no-cpu > gpurun_out/r2w_bench.json 2>gpurun_out/r2w_bench.err; cut -c1-400 gpurun_out/r2w_bench.json
MJB_SPLIT=1 timeout 600 ncu --section SourceCounters --section LaunchStats --section Occupancy --import-source on --clock-control none -k regex:'k_position|k_collision|k_constraint|k_velocity' -s 800 -c 4 -o gpurun_out/r2w_k4 -f python tools/prof_step.py 205 8192 > gpurun_out/r2w_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2w_*
