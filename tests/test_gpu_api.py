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


@pytest.mark.parametrize("cone", ["pyramidal", "elliptic"])
def test_contact_force_matches_decode(built, cone):
  """contact_force (reference support.py:326-442): pyramid / elliptic decode of efc.force per contact, contact and world frame;
  the normal forces of the humanoid's foot contacts carry its weight once it has settled for a few steps."""
  import mujoco_warp_b200 as mjw
  from mujoco_warp_b200._src import constants as C

  mjm = mjw.mjcf.load_any(util.HUMANOID)
  if cone == "elliptic":
    mjm.opt.cone = C.CONE_ELLIPTIC
  m = mjw.put_model(mjm)
  nworld = 4
  d = mjw.make_data(mjm, nworld=nworld, nconmax=24, njmax=128, m=m)
  mjw.reset_data_keyframe(m, d, 0)
  for _ in range(3):
    mjw.step(m, d)
  mjw.forward(m, d)
  nacon = int(d.nacon.cpu()[0])
  assert nacon > 0
  ids = torch.arange(nacon + 3, dtype=torch.int32, device="cuda")  # three ids past the pool end: left untouched
  out = torch.full((nacon + 3, 6), -7.0, device="cuda")
  outw = torch.full((nacon + 3, 6), -7.0, device="cuda")
  mjw.contact_force(m, d, ids, False, out)
  mjw.contact_force(m, d, ids, True, outw)
  torch.cuda.synchronize()
  out, outw = out.cpu().numpy(), outw.cpu().numpy()
  assert (out[nacon:] == -7.0).all()
  force = d.efc.force.cpu().numpy()
  adr = d.contact.efc_address[:nacon].cpu().numpy()
  dim = d.contact.dim[:nacon].cpu().numpy()
  wid = d.contact.worldid[:nacon].cpu().numpy()
  fri = d.contact.friction[:nacon].cpu().numpy()
  frame = d.contact.frame[:nacon].cpu().numpy().reshape(nacon, 3, 3)
  want = np.zeros((nacon, 6))
  for c in range(nacon):
    a0 = adr[c, 0]
    if a0 < 0:
      continue
    f = force[wid[c]]
    if cone == "pyramidal":
      if dim[c] == 1:
        want[c, 0] = f[a0]
      else:
        for i in range(dim[c] - 1):
          d1, d2 = f[a0 + 2 * i], f[a0 + 2 * i + 1]
          want[c, 0] += d1 + d2
          want[c, i + 1] = (d1 - d2) * fri[c, i]
    else:
      for i in range(dim[c]):
        want[c, i] = f[adr[c, i]]
  np.testing.assert_allclose(out[:nacon], want, rtol=1e-6, atol=1e-6)
  wantw = np.concatenate([np.einsum("ci,cik->ck", want[:, :3], frame), np.einsum("ci,cik->ck", want[:, 3:], frame)], axis=1)
  np.testing.assert_allclose(outw[:nacon], wantw, rtol=1e-5, atol=1e-5)
  assert (out[:nacon, 0] >= 0).all() and out[:nacon, 0].sum() > 0  # normal forces push


def test_get_set_state_roundtrip(built):
  import mujoco_warp_b200 as mjw

  mjm = mjw.mjcf.load_string(util.EQUALITY_XML)
  m = mjw.put_model(mjm)
  nworld = 3
  d = mjw.make_data(mjm, nworld=nworld, nconmax=16, njmax=64, m=m)
  g = torch.Generator(device="cuda").manual_seed(0)
  d.qvel.copy_(torch.randn(d.qvel.shape, device="cuda", generator=g))
  d.ctrl.copy_(torch.randn(d.ctrl.shape, device="cuda", generator=g))
  mjw.step(m, d)
  sig = int(mjw.State.INTEGRATION)
  size = 1 + mjm.nq + 2 * mjm.nv + mjm.nu + mjm.nv + 6 * mjm.nbody + mjm.neq + 7 * mjm.nmocap
  state = torch.zeros((nworld, size), device="cuda")
  mjw.get_state(m, d, state, sig)
  s = state.cpu().numpy()
  np.testing.assert_array_equal(s[:, 0], d.time.cpu().numpy())
  np.testing.assert_array_equal(s[:, 1 : 1 + mjm.nq], d.qpos.cpu().numpy())
  np.testing.assert_array_equal(s[:, 1 + mjm.nq : 1 + mjm.nq + mjm.nv], d.qvel.cpu().numpy())
  np.testing.assert_array_equal(s[:, 1 + mjm.nq + mjm.nv : 1 + mjm.nq + 2 * mjm.nv], d.qacc_warmstart.cpu().numpy())  # ACT, HISTORY are empty
  ref = {k: getattr(d, k).clone() for k in ("time", "qpos", "qvel", "qacc_warmstart", "ctrl", "mocap_pos", "mocap_quat")}
  for _ in range(3):
    mjw.step(m, d)
  active = torch.tensor([True, False, True], device="cuda")
  moved = d.qpos.clone()
  mjw.set_state(m, d, state, sig, active)
  for k, v in ref.items():
    got = getattr(d, k)
    assert torch.equal(got[0], v[0]) and torch.equal(got[2], v[2]), k
  assert torch.equal(d.qpos[1], moved[1])  # the inactive world keeps its state
  with pytest.raises(ValueError):
    mjw.get_state(m, d, state, 1 << 14)


def test_rungekutta4_equals_step(built):
  """forward() followed by rungekutta4() is step() for an RK4 model (reference forward.py:1368-1381)."""
  import mujoco_warp_b200 as mjw

  xml = util.MIXED_XML.replace('<option timestep="0.004"', '<option integrator="RK4" timestep="0.004"')
  mjm = mjw.mjcf.load_string(xml)
  m = mjw.put_model(mjm)
  da = mjw.make_data(mjm, nworld=4, nconmax=32, njmax=128, m=m)
  db = mjw.make_data(mjm, nworld=4, nconmax=32, njmax=128, m=m)
  qpos, qvel, ctrl, _ = util.seeded_state(mjm, 4, key=0, seed=5, qpos_noise=0.01, qvel_noise=0.3, ctrl_noise=1.0)
  for d in (da, db):
    d.qpos.copy_(torch.from_numpy(qpos.astype(np.float32))); d.qvel.copy_(torch.from_numpy(qvel.astype(np.float32))); d.ctrl.copy_(torch.from_numpy(ctrl.astype(np.float32)))
  for _ in range(3):
    mjw.step(m, da)
    mjw.forward(m, db); mjw.rungekutta4(m, db)
  torch.cuda.synchronize()
  assert torch.equal(da.qpos, db.qpos) and torch.equal(da.qvel, db.qvel) and torch.equal(da.time, db.time)


def test_sparse_models_expose_the_reference_csr_jacobian(built):
  """Models the reference treats as sparse (nv > 32 under jacobian = auto, io.py:153): Model.is_sparse is honest and Data.efc carries
  the reference's CSR arrays (types.py:2021-2072) -- rows in efc order, row addresses the running sum of rownnz, contact rows listing the
  dof chains of the two bodies in descending order up to their first common dof (constraint.py:2728-2753), single-dof rows for dof
  friction / joint limits.  Expanding them reproduces the dense rows the solver works on."""
  import mujoco_warp_b200 as mjw
  from mujoco_warp_b200._src.mjcf import MjDataLite, reset_data_keyframe

  mjm = mjw.mjcf.load_any(util.THREE_HUMANOIDS)
  m = mjw.put_model(mjm)
  assert m.is_sparse and mjm.nv > 32
  nworld = 4
  mjd = MjDataLite(mjm)
  reset_data_keyframe(mjm, mjd, 0)
  d = mjw.put_data(mjm, mjd, nworld=nworld, nconmax=100, njmax=192, m=m)
  assert d.efc.J.shape == (nworld, 1, d.njmax_nnz) and d.efc.J_rownnz.shape == (nworld, 192) and d.efc.J_colind.shape == (nworld, 1, d.njmax_nnz)
  for _ in range(3):
    mjw.step(m, d)
  mjw.forward(m, d)
  torch.cuda.synchronize()
  assert not (d.overflow.cpu().numpy() & int(mjw.OverflowType.NJMAX_NNZ)).any()
  nefc = d.nefc.cpu().numpy()
  assert nefc.min() > 8
  dense = d.efc.J_dense.cpu().numpy()
  np.testing.assert_array_equal(util.dense_J(d)[:, :, : mjm.nv] != 0, (dense[:, :, : mjm.nv] != 0) & (np.arange(192)[None, :, None] < nefc[:, None, None]))
  np.testing.assert_array_equal(util.dense_J(d)[0, : nefc[0]], dense[0, : nefc[0]])
  nnz, adr, col = d.efc.J_rownnz.cpu().numpy(), d.efc.J_rowadr.cpu().numpy(), d.efc.J_colind.cpu().numpy()[:, 0]
  typ = d.efc.type.cpu().numpy()
  dpar = np.asarray(mjm.dof_parentid)
  for w in range(nworld):
    ne = int(nefc[w])
    np.testing.assert_array_equal(adr[w, :ne], np.concatenate([[0], np.cumsum(nnz[w, : ne - 1])]))  # running sum in row order
    for r in range(ne):
      c = col[w, adr[w, r] : adr[w, r] + nnz[w, r]]
      if typ[w, r] in (int(mjw.ConstraintType.FRICTION_DOF), int(mjw.ConstraintType.LIMIT_JOINT)):
        assert len(c) in (1, 3)
      else:  # contact row: strictly descending dofs, every listed dof's parent chain stays inside one of at most two chains
        assert (np.diff(c) < 0).all() and len(c) >= 1
        heads = [x for x in c if not any(dpar[y] == x for y in c)]
        assert 1 <= len(heads) <= 2


def test_dense_models_keep_the_dense_layout(built):
  import mujoco_warp_b200 as mjw

  mjm = mjw.mjcf.load_any(util.HUMANOID)
  m = mjw.put_model(mjm)
  d = mjw.make_data(mjm, nworld=2, nconmax=24, njmax=64, m=m)
  assert not m.is_sparse and d.njmax_nnz == 0 and d.efc.J.shape == (2, 64, m.nv_pad) and d.efc.J_rownnz.shape == (2, 0) and d.efc.J_colind.shape == (2, 0, 0)


def test_event_trace_has_the_reference_keys_and_the_traced_step_is_the_fused_step(built):
  """testspeed --event_trace: nested keys of the reference's EventTracer (warp_util.py:51-145 flattened by testspeed.py:75-89); the
  stage-wise step that produces them must leave exactly the state the fused step leaves."""
  import mujoco_warp_b200 as mjw

  mjm = mjw.mjcf.load_any(util.HUMANOID)
  m = mjw.put_model(mjm)
  qpos, qvel, ctrl, warm = util.seeded_state(mjm, 16, key=0, seed=3)
  f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
  ds = []
  for _ in range(2):
    d = mjw.make_data(mjm, nworld=16, nconmax=24, njmax=64, m=m)
    d.qpos.copy_(f32(qpos)); d.qvel.copy_(f32(qvel)); d.ctrl.copy_(f32(ctrl)); d.qacc_warmstart.copy_(f32(warm))
    ds.append(d)
  tr = mjw.flatten_trace(mjw.event_trace_step(m, ds[0]))
  mjw.step(m, ds[1])
  torch.cuda.synchronize()
  for k in ("step", "step.forward", "step.forward.fwd_position", "step.forward.fwd_position.fwd_kinematics.kinematics", "step.forward.fwd_position.fwd_kinematics.com_pos",
            "step.forward.fwd_position.crb", "step.forward.fwd_position.collision", "step.forward.fwd_position.make_constraint", "step.forward.fwd_position.transmission",
            "step.forward.fwd_velocity", "step.forward.fwd_actuation", "step.forward.fwd_acceleration", "step.forward.solve", "step.euler"):
    assert k in tr and tr[k] >= 0.0, k
  assert tr["step"] >= tr["step.forward"] >= tr["step.forward.fwd_position"]
  for f in ("qpos", "qvel", "qacc", "qacc_warmstart"):
    np.testing.assert_array_equal(getattr(ds[0], f).cpu().numpy(), getattr(ds[1], f).cpu().numpy(), err_msg=f)


def test_override_model_reaches_the_kernels(built):
  import mujoco_warp_b200 as mjw

  mjm = mjw.mjcf.load_any(util.HUMANOID)
  m = mjw.put_model(mjm)
  d = mjw.make_data(mjm, nworld=4, nconmax=24, njmax=64, m=m)
  from mujoco_warp_b200._src.mjcf import MjDataLite, reset_data_keyframe
  mjd = MjDataLite(mjm); reset_data_keyframe(mjm, mjd, 0)
  d = mjw.put_data(mjm, mjd, nworld=4, nconmax=24, njmax=64, m=m)
  mjw.forward(m, d)
  torch.cuda.synchronize()
  assert int(d.nacon.cpu()[0]) > 0 and int(d.solver_niter.max().cpu()) >= 1
  mjw.override_model(m, ["opt.disableflags = contact", "opt.iterations = 0"])
  assert m.opt.disableflags == int(mjw.DisableBit.CONTACT) and m.opt.iterations == 0
  mjw.forward(m, d)
  torch.cuda.synchronize()
  assert int(d.nacon.cpu()[0]) == 0 and int(d.solver_niter.max().cpu()) == 0
  with pytest.raises(ValueError, match="Unrecognized model field"):
    mjw.override_model(m, {"opt.nonsense": 1})
  # direct assignment goes through the same rebinding hook (used to be a silent no-op)
  m.opt.disableflags = 0
  m.opt.iterations = 100
  m.opt.timestep = 0.001
  mjw.step(m, d)
  torch.cuda.synchronize()
  assert int(d.nacon.cpu()[0]) > 0 and abs(float(d.time[0].cpu()) - 0.001) < 1e-9
  with pytest.raises(ValueError, match="must match"):
    d.qpos = torch.zeros(3, 3, device="cuda")
  new_qpos = d.qpos.clone()
  new_qpos[:, 2] += 1.0  # lift every humanoid off the floor
  d.qpos = new_qpos
  mjw.forward(m, d)
  torch.cuda.synchronize()
  assert int(d.nacon.cpu()[0]) == 0


def test_batched_model_fields_match_per_world_models(built):
  """put_model(batch_sizes=...) (reference io.py:259-282): per-world masses, inertias, dof damping / armature, geom friction, actuator
  gains and gear.  World w must evolve exactly like a single-world run on a Model that holds w's values in its shared fields; a field
  batched with fewer entries than worlds wraps around (w % n)."""
  import mujoco_warp_b200 as mjw
  from mujoco_warp_b200._src.mjcf import MjDataLite, reset_data_keyframe

  mjm = mjw.mjcf.load_any(util.HUMANOID)
  mjm.opt.disableflags = int(mjm.opt.disableflags) & ~int(mjw.DisableBit.EULERDAMP)  # make dof_damping matter in the integrator too
  nworld = 6
  fields = {"body_mass": nworld, "body_inertia": nworld, "dof_damping": nworld, "dof_armature": 3, "geom_friction": nworld, "actuator_gainprm": 2, "actuator_gear": nworld}
  m = mjw.put_model(mjm, batch_sizes=fields)
  with pytest.raises(ValueError, match="not a batched array field"):
    mjw.put_model(mjm, batch_sizes={"body_parentid": 2})
  rng = np.random.default_rng(0)
  scale = {}
  for f, n in fields.items():
    t = getattr(m, f)
    assert t.shape[0] == n
    sc = torch.from_numpy(rng.uniform(0.7, 1.4, size=(n,) + (1,) * (t.dim() - 1)).astype(np.float32)).cuda()
    scale[f] = sc.cpu().numpy()
    t.mul_(sc)  # in place: the kernels read the registered buffer
  mjd = MjDataLite(mjm)
  reset_data_keyframe(mjm, mjd, 0)
  d = mjw.put_data(mjm, mjd, nworld=nworld, nconmax=24, njmax=64, m=m)
  ctrl = torch.from_numpy(rng.uniform(-0.5, 0.5, (nworld, mjm.nu)).astype(np.float32)).cuda()
  d.ctrl.copy_(ctrl)
  for _ in range(5):
    mjw.step(m, d)
  torch.cuda.synchronize()
  batched = {f: getattr(d, f).cpu().numpy() for f in ("qpos", "qvel", "qacc", "qfrc_bias", "actuator_force", "nefc")}
  for w in range(nworld):
    m1 = mjw.put_model(mjm)
    for f, n in fields.items():
      getattr(m1, f).mul_(torch.from_numpy(scale[f][w % n : w % n + 1]).cuda())
    d1 = mjw.put_data(mjm, mjd, nworld=1, nconmax=24, njmax=64, m=m1)
    d1.ctrl.copy_(ctrl[w : w + 1])
    for _ in range(5):
      mjw.step(m1, d1)
    torch.cuda.synchronize()
    for f, got in batched.items():
      np.testing.assert_allclose(got[w], getattr(d1, f).cpu().numpy()[0], rtol=2e-5, atol=2e-6, err_msg=f"world {w}: {f}")
  assert np.abs(batched["qpos"] - batched["qpos"][0]).max() > 1e-4  # the worlds really differ
  # assigning a tensor with another leading size re-registers the field (w % size)
  m.dof_damping = m.dof_damping[:2].clone()
  mjw.step(m, d)
  torch.cuda.synchronize()
  assert not np.isnan(d.qpos.cpu().numpy()).any()


def test_stateful_actuators_state_keyframe_and_analytic_response(built):
  """na > 0 through the public API: the ACT component of get_state / set_state, keyframe activations, the activation limit of an
  integrator, the closed-form response of an exact filter (act -> ctrl with time constant tau, forward.py:135-218 / support.py:38),
  and per-world (batched) dynprm."""
  import mujoco_warp_b200 as mjw

  xml = util.actuators_xml().replace('<key name="k0" qpos="0.3 -0.5 0.2 0.01 0 0 0.1"/>', '<key name="k0" qpos="0.3 -0.5 0.2 0.01 0 0 0.1" act="0.1 0.2 0.3 0.02 -0.1 0.05"/>')
  mjm = mjw.mjcf.load_string(xml)
  assert mjm.na == 6
  nworld = 4
  m = mjw.put_model(mjm, batch_sizes={"actuator_dynprm": nworld})
  tau = torch.tensor([0.02, 0.04, 0.08, 0.16], device="cuda")
  dp = m.actuator_dynprm.clone()
  dp[:, 2, 0] = tau  # a_fex (filterexact): one time constant per world
  m.actuator_dynprm = dp
  d = mjw.make_data(mjm, nworld=nworld, nconmax=16, njmax=64, m=m)
  mjw.reset_data_keyframe(m, d, 0)
  np.testing.assert_allclose(d.act.cpu().numpy(), np.tile([0.1, 0.2, 0.3, 0.02, -0.1, 0.05], (nworld, 1)), rtol=1e-6)
  ctrl = torch.zeros_like(d.ctrl)
  ctrl[:, 0] = 2.0   # a_int: integrates ctrl, clamped at actrange 0.6
  ctrl[:, 2] = 1.0   # a_fex
  d.ctrl.copy_(ctrl)
  nstep, dt = 100, float(mjm.opt.timestep)
  for _ in range(nstep):
    mjw.step(m, d)
  torch.cuda.synchronize()
  act = d.act.cpu().numpy()
  np.testing.assert_allclose(act[:, 0], 0.6, rtol=1e-6)  # 0.1 + 100 * 0.004 * 2 = 0.9 without the activation limit
  want = 1.0 + (0.3 - 1.0) * np.exp(-nstep * dt / tau.cpu().numpy())
  np.testing.assert_allclose(act[:, 2], want, rtol=2e-5, atol=2e-6)
  assert not np.isnan(d.qpos.cpu().numpy()).any()
  # ACT is part of the integration state
  sig = int(mjw.State.INTEGRATION)
  size = 1 + mjm.nq + 2 * mjm.nv + mjm.na + mjm.nu + mjm.nv + 6 * mjm.nbody
  state = torch.zeros((nworld, size), device="cuda")
  mjw.get_state(m, d, state, sig)
  off = 1 + mjm.nq + mjm.nv
  np.testing.assert_array_equal(state[:, off : off + mjm.na].cpu().numpy(), act)
  saved = {k: getattr(d, k).clone() for k in ("qpos", "qvel", "act", "time")}
  for _ in range(5):
    mjw.step(m, d)
  after = {k: getattr(d, k).clone() for k in saved}
  mjw.set_state(m, d, state, sig)
  for k, v in saved.items():
    assert torch.equal(getattr(d, k), v), k
  d.qacc_warmstart.copy_(state[:, 1 + mjm.nq + mjm.nv + mjm.na : 1 + mjm.nq + 2 * mjm.nv + mjm.na])
  for _ in range(5):
    mjw.step(m, d)
  for k, v in after.items():  # the restored state replays the same trajectory bit for bit
    assert torch.equal(getattr(d, k), v), k
