#!/bin/bash
# Round-end GPU pass: tests, smoke, bench lines (all workloads), ncu launch list of the bench command, ncu --set full of every kernel.
cd "${GRAFT_REPO_ROOT:-.}"
R=${R:-r02}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${R}_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${R}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
MJB_E2E_TRACE=1 timeout 600 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/${R}_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err; tail -c 600 gpurun_out/${R}_bench_reference.json
for wl in g1 three_humanoids convex_mesh; do
  timeout 400 python bench.py --workload $wl --steps 100 --warmup 20 > gpurun_out/${R}_bench_$wl.json 2>> gpurun_out/${R}_bench.err; echo "$wl rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --no-graph --steps 8 --warmup 3 --no-cpu > gpurun_out/${R}_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
MJB_SPLIT=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'k_position|k_collision|k_constraint|k_velocity|k_solver|k_euler' -s 720 -c 6 -o gpurun_out/${R}_prof -f python tools/prof_step.py 125 8192 > gpurun_out/${R}_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/${R}_prof.ncu-rep --page raw --csv > gpurun_out/${R}_raw.csv 2>/dev/null
ncu -i gpurun_out/${R}_prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${R}_src.csv 2>/dev/null
ls -la gpurun_out/${R}_*
