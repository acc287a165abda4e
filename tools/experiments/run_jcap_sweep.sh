# GPU-side A/B helper: staged Jacobian rows (MJB_JCAP) for the nv > 32 solver
for j in ${JCAP_LIST:-48 24 8 0}; do
  echo "== three JCAP=$j"; MJB_JCAP=$j python -m mujoco_warp_b200.testspeed mujoco_warp_b200/test_data/three_humanoids.npz --nworld 8192 --nconmax 100 --njmax 192 --nstep 200 2>&1 | grep -E "steps per second"
  echo "== G1 JCAP=$j"; MJB_JCAP=$j python -m mujoco_warp_b200.testspeed mujoco_warp_b200/test_data/unitree_g1_flat.npz --nworld 4096 --nconmax 48 --njmax 192 --nstep 250 --replay mujoco_warp_b200/test_data/unitree_g1_shuffle_dance.npz 2>&1 | grep -E "steps per second"
done
