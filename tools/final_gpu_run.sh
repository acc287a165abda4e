# Round-end GPU pass: parity tests, bench line, ncu launch list of the same bench command, G1 / three_humanoids throughput.
set -o pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 2500 gpurun_out/bench_final.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --no-graph --steps 8 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
echo "== G1"; python -m mujoco_warp_b200.testspeed mujoco_warp_b200/test_data/unitree_g1_flat.npz --nworld 4096 --nconmax 48 --njmax 192 --nstep 250 --replay mujoco_warp_b200/test_data/unitree_g1_shuffle_dance.npz --event_trace true 2>&1 | grep -E "steps per second|time per step|solve:|step:|fwd_|collision|constraint|euler"
echo "== three"; python -m mujoco_warp_b200.testspeed mujoco_warp_b200/test_data/three_humanoids.npz --nworld 8192 --nconmax 100 --njmax 192 --nstep 200 --event_trace true 2>&1 | grep -E "steps per second|time per step|solve:|step:"
