# GPU-side A/B helper: parity tests, then unitree G1 and three_humanoids throughput for each solver team size
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for nw in ${NW_LIST:-1}; do
  echo "== G1 NW=$nw"; MJB_SOLVER_WARPS=$nw python -m mujoco_warp_b200.testspeed mujoco_warp_b200/test_data/unitree_g1_flat.npz --nworld 4096 --nconmax 48 --njmax 192 --nstep 250 --replay mujoco_warp_b200/test_data/unitree_g1_shuffle_dance.npz --event_trace true 2>&1 | grep -E "steps per second|solve:|step:"
  echo "== three NW=$nw"; MJB_SOLVER_WARPS=$nw python -m mujoco_warp_b200.testspeed mujoco_warp_b200/test_data/three_humanoids.npz --nworld 8192 --nconmax 100 --njmax 192 --nstep 200 --event_trace true 2>&1 | grep -E "steps per second|solve:|step:"
done
