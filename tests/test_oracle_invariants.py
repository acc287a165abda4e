"""Physical invariants that pin the oracle's dynamics (the reference's own dynamics tests delegate to live MuJoCo, which is
unavailable here -- see oracle.h).  Each invariant is a property the true algorithm must satisfy independent of any
implementation: symmetric positive-definite M equal to the kinetic-energy Hessian, inverse-dynamics consistency, KKT
optimality of the constraint solve, momentum/energy behaviour of a free-falling humanoid."""

import numpy as np
import pytest

from tests import util


@pytest.fixture(scope="module")
def scene(built):
  from mujoco_warp_b200._src import mjcf

  return mjcf.load_any(util.HUMANOID)


def dense_M(mjm, Mcsr):
  M = np.zeros((mjm.nv, mjm.nv))
  for i in range(mjm.nv):
    for k in range(mjm.M_rownnz[i]):
      j = mjm.M_colind[mjm.M_rowadr[i] + k]
      M[i, j] = M[j, i] = Mcsr[mjm.M_rowadr[i] + k]
  return M


def test_M_is_spd_and_matches_kinetic_energy(scene):
  mjm = scene
  nw = 4
  o = util.make_oracle(mjm, nw, 24, 64)
  qpos, qvel, ctrl, warm = util.seeded_state(mjm, nw)
  o.set_state(qpos=qpos, qvel=qvel)
  o.forward()
  for w in range(nw):
    M = dense_M(mjm, o.d["M"][w])
    assert np.linalg.eigvalsh(M).min() > 0
    # kinetic energy two ways: 0.5 v'Mv  ==  sum_b 0.5 cvel' I_c cvel (cinert in the com frame)
    v = o.d["qvel"][w]
    ke_joint = 0.5 * v @ M @ v - 0.5 * np.sum(mjm.dof_armature * v * v)
    ke_body = 0.0
    for b in range(1, mjm.nbody):
      ci, cv = o.d["cinert"][w, b], o.d["cvel"][w, b]
      I = np.array([[ci[0], ci[3], ci[4]], [ci[3], ci[1], ci[5]], [ci[4], ci[5], ci[2]]])
      mc, mass = ci[6:9], ci[9]
      wv, lv = cv[:3], cv[3:]
      ke_body += 0.5 * (wv @ I @ wv) + lv @ np.cross(wv, mc) * 1.0 + 0.5 * mass * lv @ lv
    assert ke_joint == pytest.approx(ke_body, rel=1e-9, abs=1e-10)
    # qLD: M = U^T U
    U = o.d["qLD"][w].reshape(mjm.nv, mjm.nv)
    np.testing.assert_allclose(U.T @ U, M, atol=1e-9)


def test_smooth_dynamics_consistency(scene):
  """M qacc_smooth + qfrc_bias == qfrc_passive + qfrc_actuator + qfrc_applied  (forward.py:1255-1324)."""
  mjm = scene
  nw = 4
  o = util.make_oracle(mjm, nw, 24, 64)
  qpos, qvel, ctrl, warm = util.seeded_state(mjm, nw)
  o.set_state(qpos=qpos, qvel=qvel, ctrl=ctrl)
  o.forward()
  for w in range(nw):
    M = dense_M(mjm, o.d["M"][w])
    lhs = M @ o.d["qacc_smooth"][w] + o.d["qfrc_bias"][w]
    rhs = o.d["qfrc_passive"][w] + o.d["qfrc_actuator"][w]
    np.testing.assert_allclose(lhs, rhs, atol=1e-8, rtol=1e-9)
    np.testing.assert_allclose(o.d["qfrc_actuator"][w][6:], (np.clip(ctrl[w], -1, 1) * mjm.actuator_gear[:, 0])[np.argsort(mjm.jnt_dofadr[mjm.actuator_trnid[:, 0]])], atol=1e-12)


def test_bias_force_is_gravity_for_static_pose(scene):
  """With zero velocity, qfrc_bias = -J_com^T m g summed over bodies: the root's vertical translational dof carries total weight."""
  mjm = scene
  o = util.make_oracle(mjm, 1, 24, 64)
  o.set_state(qpos=mjm.key_qpos[2])  # no_efc keyframe
  o.forward()
  total_mass = mjm.body_mass.sum()
  assert o.d["qfrc_bias"][0, 2] == pytest.approx(total_mass * 9.81, rel=1e-12)
  assert o.d["nefc"][0] == 0 and o.d["ncon"][0] == 0
  # free fall: root linear acceleration = gravity, solver leaves qacc = qacc_smooth
  np.testing.assert_allclose(o.d["qacc"][0], o.d["qacc_smooth"][0], atol=1e-9)
  assert o.d["qacc_smooth"][0, 2] == pytest.approx(-9.81, abs=0.5)  # joint springs couple slightly into the root


def test_solver_kkt(scene):
  """At the solution of the convex problem the gradient vanishes: M qacc - qfrc_smooth - J^T f = 0, with per-row forces
  consistent with the row states (solver.py:425-477) and non-negative normal cone combinations."""
  mjm = scene
  nw = 8
  o = util.make_oracle(mjm, nw, 24, 64)
  qpos, qvel, ctrl, warm = util.seeded_state(mjm, nw)
  o.set_state(qpos=qpos, qvel=qvel, ctrl=ctrl, qacc_warmstart=warm)
  o.forward()
  assert (o.d["overflow"] == 0).all()
  for w in range(nw):
    ne = o.d["nefc"][w]
    assert ne > 0
    M = dense_M(mjm, o.d["M"][w])
    J, f = o.d["efc_J"][w, :ne], o.d["efc_force"][w, :ne]
    grad = M @ o.d["qacc"][w] - o.d["qfrc_smooth"][w] - J.T @ f
    scale = mjm.stat.meaninertia * mjm.nv
    assert np.linalg.norm(grad) / scale < 1e-6
    np.testing.assert_allclose(o.d["qfrc_constraint"][w], J.T @ f, atol=1e-9)
    np.testing.assert_allclose(o.d["efc_Ma"][w], M @ o.d["qacc"][w], atol=1e-7)
    jar = J @ o.d["qacc"][w] - o.d["efc_aref"][w, :ne]
    st = o.d["efc_state"][w, :ne]
    assert ((st == 1) == (jar < 0)).all()
    np.testing.assert_allclose(f, np.where(jar < 0, -o.d["efc_D"][w, :ne] * jar, 0.0), atol=1e-9)
    assert (f >= 0).all()


def test_contacts_squat_keyframe(scene):
  """Squat keyframe: 4 foot capsules on the floor, 2 plane-capsule contacts each, condim 3 -> 4 pyramid rows per contact."""
  mjm = scene
  o = util.make_oracle(mjm, 1, 24, 64)
  o.set_state(qpos=mjm.key_qpos[0])
  o.forward()
  assert o.d["ncon"][0] == 8 and o.d["nefc"][0] == 32 and o.d["nl"][0] == 0
  g = o.d["con_geom"][0, :8]
  assert (g[:, 0] == 0).all()
  assert sorted(set(g[:, 1])) == sorted(mjm.names.geom.index(n) for n in ("foot1_right", "foot2_right", "foot1_left", "foot2_left"))
  assert (o.d["con_dim"][0, :8] == 3).all()
  np.testing.assert_allclose(o.d["con_frame"][0, :8, 0], [[0, 0, 1]] * 8, atol=1e-12)
  assert (o.d["efc_type"][0, :32] == 6).all()


def test_free_fall_conserves_momentum_and_energy(scene):
  """no_efc keyframe, gravity only (no damping on the root): horizontal momentum stays 0, the CoM falls as g t^2 / 2."""
  mjm = scene
  o = util.make_oracle(mjm, 1, 24, 64)
  o.set_state(qpos=mjm.key_qpos[2])
  o.forward()
  com0 = o.d["subtree_com"][0, 1].copy()
  n = 40
  for _ in range(n):
    o.step()
  o.forward()
  t = n * mjm.opt.timestep
  com = o.d["subtree_com"][0, 1]
  np.testing.assert_allclose(com[:2], com0[:2], atol=1e-9)
  # semi-implicit Euler: z(t) = z0 - g dt^2 n(n+1)/2
  assert com[2] - com0[2] == pytest.approx(-9.81 * mjm.opt.timestep**2 * n * (n + 1) / 2, rel=1e-6)
  assert o.d["time"][0] == pytest.approx(t)


def test_fp32_oracle_tracks_fp64(scene):
  mjm = scene
  nw = 4
  o64 = util.make_oracle(mjm, nw, 24, 64, dtype=np.float64)
  o32 = util.make_oracle(mjm, nw, 24, 64, dtype=np.float32)
  qpos, qvel, ctrl, warm = util.seeded_state(mjm, nw)
  for o in (o64, o32):
    o.set_state(qpos=qpos.astype(np.float32), qvel=qvel.astype(np.float32), ctrl=ctrl.astype(np.float32), qacc_warmstart=warm.astype(np.float32))
    o.forward()
  np.testing.assert_array_equal(o32.d["nefc"], o64.d["nefc"])
  np.testing.assert_allclose(o32.d["qacc"], o64.d["qacc"], atol=5e-3 * np.abs(o64.d["qacc"]).max())


def _elliptic_cost(mjm, od, w, qacc):
  """Convex objective of the elliptic-cone problem written from its definition (Gauss term + per-contact cone penalty:
  0 in the top zone, full quadratic in the bottom zone, 0.5 dm (N - mu T)^2 in between)."""
  ne = od["nefc"][w]
  M = dense_M(mjm, od["M"][w])
  J, D, aref = od["efc_J"][w, :ne], od["efc_D"][w, :ne], od["efc_aref"][w, :ne]
  typ, eid = od["efc_type"][w, :ne], od["efc_id"][w, :ne]
  jar = J @ qacc - aref
  dq = qacc - od["qacc_smooth"][w]
  cost = 0.5 * dq @ M @ dq
  done = set()
  for r in range(ne):
    if typ[r] != 7:  # limit / frictionless contact rows
      cost += 0.5 * D[r] * jar[r] ** 2 if jar[r] < 0 else 0.0
      continue
    c = eid[r]
    if c in done:
      continue
    done.add(c)
    dim, fri = od["con_dim"][w, c], od["con_friction"][w, c]
    mu = fri[0] / np.sqrt(mjm.opt.impratio)
    rows = np.arange(r, r + dim)
    N = jar[r] * mu
    T = np.sqrt(((jar[rows[1:]] * fri[: dim - 1]) ** 2).sum())
    if N >= mu * T:
      continue
    if mu * N + T <= 0:
      cost += 0.5 * (D[rows] * jar[rows] ** 2).sum()
    else:
      cost += 0.5 * D[r] / (mu * mu * (1 + mu * mu)) * (N - mu * T) ** 2
  return cost


def test_solver_elliptic_minimises_cone_cost(scene):
  """Elliptic cones: the oracle's Newton solution is the minimiser of the independently written convex objective (finite-
  difference gradient ~ 0, random perturbations never decrease it), forces obey the friction cone, and the CONE zone is hit."""
  import copy

  from mujoco_warp_b200._src import constants as C

  mjm = copy.deepcopy(scene)
  mjm.opt.cone = C.CONE_ELLIPTIC
  mjm.opt.tolerance = 1e-10
  nw = 8
  o = util.make_oracle(mjm, nw, 24, 64, clamp_tolerance=False)
  qpos, qvel, ctrl, warm = util.seeded_state(mjm, nw, seed=5)
  o.set_state(qpos=qpos, qvel=qvel, ctrl=ctrl, qacc_warmstart=warm)
  o.forward()
  od = o.d
  assert (od["overflow"] == 0).all()
  assert (od["efc_state"] == 4).any() and (od["efc_type"] == 7).any()
  rng = np.random.default_rng(0)
  for w in range(nw):
    ne = od["nefc"][w]
    if ne == 0:
      continue
    M = dense_M(mjm, od["M"][w])
    J, f = od["efc_J"][w, :ne], od["efc_force"][w, :ne]
    grad = M @ od["qacc"][w] - od["qfrc_smooth"][w] - J.T @ f
    scale = mjm.stat.meaninertia * mjm.nv
    assert np.linalg.norm(grad) / scale < 1e-6
    q = od["qacc"][w].copy()
    c0 = _elliptic_cost(mjm, od, w, q)
    # the analytic force is minus the cost gradient wrt Jaref: check by central differences along random directions
    for _ in range(6):
      dq = rng.standard_normal(mjm.nv)
      h = 1e-4
      num = (_elliptic_cost(mjm, od, w, q + h * dq) - _elliptic_cost(mjm, od, w, q - h * dq)) / (2 * h)
      assert abs(num) < 1e-4 * max(1.0, abs(c0)), (w, num, c0)
      assert _elliptic_cost(mjm, od, w, q + 1e-2 * dq) >= c0 - 1e-9 * max(1.0, abs(c0))
    # friction cone on the contact forces (impratio = 1): f_n >= 0 and sum (f_j / mu_j)^2 <= f_n^2
    typ, eid = od["efc_type"][w, :ne], od["efc_id"][w, :ne]
    for r in range(ne):
      if typ[r] == 7 and (r == 0 or eid[r - 1] != eid[r] or typ[r - 1] != 7):
        c = eid[r]
        dim, fri = od["con_dim"][w, c], od["con_friction"][w, c]
        ft = f[r + 1 : r + dim] / fri[: dim - 1]
        assert f[r] >= -1e-9
        if od["efc_state"][w, r] == 4:
          np.testing.assert_allclose(np.sqrt((ft**2).sum()), f[r], rtol=1e-6, atol=1e-9)


SENSOR_BALL = """
<mujoco>
  <option timestep="0.002" gravity="0 0 -9.81"/>
  <worldbody>
    <geom name="floor" type="plane" size="0 0 .05"/>
    <body name="ball" pos="0 0 {z}">
      <freejoint/>
      <geom type="sphere" size="0.1" mass="2"/>
      <site name="s" pos="0.02 0.01 0" euler="30 20 10"/>
      <site name="zone" type="box" size="0.2 0.2 0.2"/>
    </body>
  </worldbody>
  <sensor>
    <accelerometer name="acc" site="s"/> <gyro name="gyro" site="s"/> <velocimeter name="vel" site="s"/>
    <touch name="touch" site="zone"/> <force name="frc" site="s"/>
    <subtreelinvel name="lv" body="ball"/> <subtreeangmom name="am" body="ball"/> <framelinacc name="fla" objtype="site" objname="s"/>
  </sensor>
</mujoco>"""


def _sens(mjm, od, name, w=0):
  i = mjm.names.sensor.index(name)
  return od["sensordata"][w, mjm.sensor_adr[i] : mjm.sensor_adr[i] + mjm.sensor_dim[i]]


def test_sensor_physics_resting_and_free_fall(built):
  """Implementation-independent facts about the sensors: a body resting on the floor has an accelerometer (and frame-acceleration) reading of +g along the
  world vertical, a touch sensor reading its weight; in free fall both read the centripetal term of the offset site only (no gravity); the velocity sensors follow the rigid-body velocity field."""
  from mujoco_warp_b200._src import mjcf

  # resting: let the contact settle, then read
  mjm = mjcf.load_string(SENSOR_BALL.format(z=0.0995))
  o = util.make_oracle(mjm, 1, 8, 32, clamp_tolerance=False)
  for _ in range(400):
    o.step()
  o.forward()
  kin = mjcf.kinematics_np(mjm, o.d["qpos"][0])
  R = np.asarray(kin.site_xmat).reshape(-1, 3, 3)[mjm.names.site.index("s")]
  np.testing.assert_allclose(R @ _sens(mjm, o.d, "acc"), [0, 0, 9.81], atol=2e-3)
  np.testing.assert_allclose(_sens(mjm, o.d, "touch"), [2 * 9.81], rtol=2e-3)
  np.testing.assert_allclose(_sens(mjm, o.d, "fla"), [0, 0, 9.81], atol=2e-3)  # like the accelerometer (MuJoCo's cacc carries -gravity), world frame
  # interaction force the body exerts on its parent (the world) through its joint: zero for a free joint
  np.testing.assert_allclose(_sens(mjm, o.d, "frc"), [0, 0, 0], atol=2e-2)

  # free fall with spin: accelerometer 0, frame acceleration = gravity + centripetal term, velocity field of a rigid body
  mjm = mjcf.load_string(SENSOR_BALL.format(z=2.0))
  o = util.make_oracle(mjm, 1, 8, 32)
  wvec, vvec = np.array([0.3, -0.2, 0.5]), np.array([0.1, 0.2, -0.3])
  qvel = np.concatenate([vvec, wvec])[None]  # free joint: linear velocity in the world frame, angular velocity in the body frame (identity here)
  o.set_state(qvel=qvel)
  o.forward()
  kin = mjcf.kinematics_np(mjm, o.d["qpos"][0])
  si = mjm.names.site.index("s")
  R, p = np.asarray(kin.site_xmat).reshape(-1, 3, 3)[si], np.asarray(kin.site_xpos).reshape(-1, 3)[si]
  r = p - o.d["qpos"][0, :3]
  np.testing.assert_allclose(R @ _sens(mjm, o.d, "acc"), np.cross(wvec, np.cross(wvec, r)), atol=1e-9)  # no gravity felt: centripetal term of the offset site only
  np.testing.assert_allclose(R @ _sens(mjm, o.d, "gyro"), wvec, atol=1e-12)
  np.testing.assert_allclose(R @ _sens(mjm, o.d, "vel"), vvec + np.cross(wvec, r), atol=1e-12)
  np.testing.assert_allclose(_sens(mjm, o.d, "fla"), np.cross(wvec, np.cross(wvec, r)), atol=1e-9)  # centripetal term only
  np.testing.assert_allclose(_sens(mjm, o.d, "lv"), vvec, atol=1e-12)
  inertia = 0.4 * 2 * 0.1**2  # solid sphere
  np.testing.assert_allclose(_sens(mjm, o.d, "am"), inertia * wvec, rtol=1e-9)
  np.testing.assert_allclose(_sens(mjm, o.d, "touch"), [0.0])
