"""Public API additions beyond step/forward: split step, single sub-stages, solve_m / mul_m, keyframe resets, host readback."""

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene(built):
  import mujoco_warp_b200 as mjw

  mjm = mjw.mjcf.load_any(util.HUMANOID)
  return mjw, mjm, mjw.put_model(mjm)


def _data(scene, nworld=8, seed=5):
  mjw, mjm, m = scene
  d = mjw.make_data(mjm, nworld=nworld, nconmax=24, njmax=64, m=m)
  qpos, qvel, ctrl, warm = util.seeded_state(mjm, nworld, seed=seed)
  for name, val in (("qpos", qpos), ("qvel", qvel), ("ctrl", ctrl), ("qacc_warmstart", warm)):
    getattr(d, name).copy_(torch.from_numpy(val.astype(np.float32)))
  return d


def test_split_step_and_substages_match_fused(scene):
  mjw, mjm, m = scene
  d1, d2, d3 = _data(scene), _data(scene), _data(scene)
  mjw.step(m, d1)
  mjw.step1(m, d2)
  mjw.step2(m, d2)
  # d3: every stage by hand, fwd_velocity split into its parts
  mjw.fwd_kinematics(m, d3); mjw.crb(m, d3); mjw.collision(m, d3); mjw.make_constraint(m, d3); mjw.transmission(m, d3)
  mjw.fwd_velocity(m, d3)  # (actuator_velocity has no separate entry point)
  ref = {n: getattr(d3, n).clone() for n in ("cvel", "cdof_dot", "qfrc_passive", "qfrc_bias", "cacc", "cfrc_int")}
  for n in ref:
    getattr(d3, n).zero_()
  mjw.com_vel(m, d3); mjw.passive(m, d3); mjw.rne(m, d3)
  for n, want in ref.items():
    assert torch.equal(getattr(d3, n), want), n
  mjw.fwd_actuation(m, d3); mjw.fwd_acceleration(m, d3); mjw.solve(m, d3); mjw.euler(m, d3)
  torch.cuda.synchronize()
  for n in ("qpos", "qvel", "qacc", "qacc_warmstart", "time"):
    assert torch.equal(getattr(d1, n), getattr(d2, n)), n
    assert torch.equal(getattr(d1, n), getattr(d3, n)), n
  with pytest.raises(NotImplementedError):
    mjw.rne(m, d3, flg_acc=True)
  with pytest.raises(NotImplementedError):
    mjw.implicit(m, d3)  # humanoid integrates with Euler


def test_solve_m_and_mul_m(scene):
  mjw, mjm, m = scene
  d = _data(scene)
  mjw.forward(m, d)
  rng = np.random.default_rng(0)
  y = torch.from_numpy(rng.standard_normal((d.nworld, m.nv)).astype(np.float32)).cuda()
  x = torch.empty_like(y)
  back = torch.empty_like(y)
  mjw.solve_m(m, d, x, y)
  mjw.mul_m(m, d, back, x)
  torch.cuda.synchronize()
  np.testing.assert_allclose(back.cpu().numpy(), y.cpu().numpy(), atol=2e-4, rtol=2e-4)
  # qacc_smooth = M^-1 qfrc_smooth
  mjw.solve_m(m, d, x, d.qfrc_smooth)
  np.testing.assert_allclose(x.cpu().numpy(), d.qacc_smooth.cpu().numpy(), atol=1e-3, rtol=1e-3)
  # dense check of mul_m against the CSR inertia of world 0
  M = np.zeros((mjm.nv, mjm.nv))
  Mw = d.M[0].cpu().numpy()
  for i in range(mjm.nv):
    for k in range(mjm.M_rownnz[i]):
      j = mjm.M_colind[mjm.M_rowadr[i] + k]
      M[i, j] = M[j, i] = Mw[mjm.M_rowadr[i] + k]
  mjw.mul_m(m, d, back, y)
  np.testing.assert_allclose(back[0].cpu().numpy(), M @ y[0].cpu().numpy(), atol=1e-4, rtol=1e-4)


def test_reset_data_keyframe_and_get_data_into(scene):
  mjw, mjm, m = scene
  from mujoco_warp_b200._src.mjcf import MjDataLite

  d = _data(scene)
  for _ in range(3):
    mjw.step(m, d)
  before = d.qpos.clone()
  keys = torch.tensor([0, 1, -1, 2, 7, 0, 99, 1], dtype=torch.int32)
  mjw.reset_data_keyframe(m, d, keys)
  torch.cuda.synchronize()
  for w, k in enumerate(keys.tolist()):
    if 0 <= k < mjm.nkey:
      np.testing.assert_allclose(d.qpos[w].cpu().numpy(), mjm.key_qpos[k].astype(np.float32))
      assert float(d.qvel[w].abs().max()) == 0.0 and float(d.time[w]) == float(mjm.key_time[k])
    else:
      assert torch.equal(d.qpos[w], before[w])
  with pytest.raises(ValueError):
    mjw.reset_data_keyframe(m, d, mjm.nkey)
  mjw.reset_data_keyframe(m, d, 0)
  mjw.forward(m, d)
  res = mjw.get_data_into(MjDataLite(mjm), mjm, d, world_id=3)
  assert res.ncon == 8 and res.nefc == 32 and res.efc_J.shape == (32, mjm.nv) and res.contact["geom"].shape == (8, 2)
  np.testing.assert_allclose(res.qpos, mjm.key_qpos[0], atol=1e-6)
  np.testing.assert_allclose(res.qacc, d.qacc[3].cpu().numpy(), atol=0)
