"""Minimal step loop for ncu captures:  ncu ... python tools/prof_step.py [nsteps] [nworld]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import mujoco_warp_b200 as mjw
from mujoco_warp_b200._src.mjcf import MjDataLite, reset_data_keyframe
from tests import util

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
nworld = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
mjm = mjw.mjcf.load_any(util.HUMANOID)
m = mjw.put_model(mjm)
mjd = MjDataLite(mjm)
reset_data_keyframe(mjm, mjd, 0)
d = mjw.put_data(mjm, mjd, nworld=nworld, nconmax=24, njmax=64, m=m)
center = torch.from_numpy(np.asarray(mjm.key_ctrl[0], dtype=np.float32)).cuda()
for i in range(nsteps):
  mjw.ctrl_noise(m, d, i, center)
  mjw.step(m, d)
torch.cuda.synchronize()
print("nefc mean", float(d.nefc.float().mean()), "niter mean", float(d.solver_niter.float().mean()))
