"""Golden trajectories from the reference's own step pipeline, executed on the CPU through tools/warp_shim.py.

  python tools/make_pipeline_goldens.py          # writes tests/golden/pipeline_<scene>.npz

For every scene the UNMODIFIED reference code (io.put_model -> io.make_data -> forward.forward / forward.step and every
kernel they launch: smooth.py, collision_driver.py, collision_primitive.py, constraint.py, passive.py, solver.py,
derivative.py) runs in double precision on a model compiled by mujoco_warp_b200._src.mjcf, from seeded states.  The fixture
stores the inputs and every Data field the parity tests compare, after forward() and after each of NSTEP step() calls.
tests/test_oracle_golden_pipeline.py holds the fp64 oracle to these numbers.
"""

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mujoco_warp_b200._src import constants as C  # noqa: E402
from mujoco_warp_b200._src import mjcf  # noqa: E402
from tests import util  # noqa: E402
from tests.test_gpu_colliders import BOX_XML  # noqa: E402
from tools import ref_runner  # noqa: E402

NSTEP = 4

FIELDS = [
  "xpos", "xquat", "xmat", "xipos", "ximat", "xanchor", "xaxis", "geom_xpos", "geom_xmat", "site_xpos", "site_xmat", "cam_xpos", "cam_xmat",
  "light_xpos", "light_xdir", "subtree_com", "cdof", "cinert", "crb", "M", "actuator_length", "actuator_moment", "actuator_velocity", "cvel",
  "cdof_dot", "qfrc_bias", "qfrc_spring", "qfrc_damper", "qfrc_gravcomp", "qfrc_passive", "actuator_force", "qfrc_actuator", "qfrc_smooth", "qacc_smooth",
  "qacc", "qfrc_constraint", "ne", "nf", "nl", "nefc", "solver_niter", "qLU", "sensordata", "subtree_linvel", "subtree_angmom", "cfrc_ext", "act_dot", "ten_length", "ten_velocity",
]
EFC = ["type", "id", "J", "pos", "margin", "D", "vel", "aref", "frictionloss", "force", "state"]
CON = ["dist", "pos", "frame", "includemargin", "friction", "solref", "solreffriction", "solimp", "dim", "geom", "efc_address", "worldid", "geomcollisionid"]


def scenes():
  hum = mjcf.load_any(util.HUMANOID)
  yield "humanoid", hum, dict(nconmax=24, njmax=128, key=0, qpos_noise=0.003, exact_world0=False)
  # capacity overflow (collision_core.py:271-291, constraint.py:2048,2712): one world lying flat on the floor detects 25 contacts / 100
  # rows but the pool holds 20 contacts (19 broadphase candidates fit) and njmax is 64 -> 5 contacts dropped, the last kept contacts lose rows, overflow bits set.
  # One world, because the contact pool is global: which WORLD loses contacts depends on the order worlds reach the atomic counter.
  yield "humanoid_trunc", mjcf.load_any(util.HUMANOID), dict(nconmax=20, njmax=64, key=None, qpos_noise=0.0, qvel_noise=0.2, ctrl_noise=0.3, exact_world0=True, nworld=1)
  hum_e = mjcf.load_any(util.HUMANOID)
  hum_e.opt.cone = C.CONE_ELLIPTIC
  yield "humanoid_elliptic", hum_e, dict(nconmax=24, njmax=128, key=0, qpos_noise=0.003, exact_world0=False)
  hum_cg = mjcf.load_any(util.HUMANOID)
  hum_cg.opt.solver = C.SOL_CG
  yield "humanoid_cg", hum_cg, dict(nconmax=24, njmax=128, key=0, qpos_noise=0.003, exact_world0=False)
  yield "mixed", mjcf.load_string(util.MIXED_XML), dict(nconmax=32, njmax=128, key=0, qpos_noise=0.01, qvel_noise=0.3, ctrl_noise=1.5, exact_world0=False)
  ell = util.MIXED_XML.replace('<option timestep="0.004"', '<option cone="elliptic" impratio="2" timestep="0.004"')
  yield "mixed_elliptic", mjcf.load_string(ell), dict(nconmax=32, njmax=128, key=0, qpos_noise=0.01, qvel_noise=0.3, ctrl_noise=1.5, exact_world0=False)
  cg = util.MIXED_XML.replace('<option timestep="0.004"', '<option solver="CG" cone="elliptic" timestep="0.004"')
  yield "mixed_elliptic_cg", mjcf.load_string(cg), dict(nconmax=32, njmax=128, key=0, qpos_noise=0.01, qvel_noise=0.3, ctrl_noise=1.5, exact_world0=False)
  yield "pairs", mjcf.load_string(util.pairs_xml()), dict(nconmax=32, njmax=128, key=0, qpos_noise=0.01, qvel_noise=0.3, ctrl_noise=1.5, exact_world0=False)
  yield "passive", mjcf.load_string(util.passive_xml()), dict(nconmax=32, njmax=128, key=0, qpos_noise=0.01, qvel_noise=0.3, ctrl_noise=1.5, exact_world0=False)
  yield "convex", mjcf.load_string(util.CONVEX_XML), dict(nconmax=64, njmax=256, key=None, qpos_noise=0.004, qvel_noise=0.05, ctrl_noise=0.0, exact_world0=False)
  yield "boxes", mjcf.load_string(BOX_XML), dict(nconmax=48, njmax=200, key=None, qpos_noise=0.003, qvel_noise=0.05, ctrl_noise=0.0, exact_world0=False)
  yield "boxccd", mjcf.load_string(util.boxccd_xml()), dict(nconmax=48, njmax=256, key=None, qpos_noise=0.0004, qvel_noise=0.05, ctrl_noise=0.0, exact_world0=False)
  yield "boxccd_mixed", mjcf.load_string(util.boxccd_xml(True)), dict(nconmax=48, njmax=256, key=None, qpos_noise=0.0004, qvel_noise=0.05, ctrl_noise=0.0, exact_world0=False)
  # sweep-and-prune broadphase (collision_driver.py:582): same scenes, tile sort with the default filters and segmented sort
  # without the bounding-sphere filter (so the sweep's own pruning decides which pairs reach the OBB test)
  sap = mjcf.load_string(util.MIXED_XML)
  sap.opt.broadphase = 1
  yield "mixed_sap", sap, dict(nconmax=32, njmax=128, key=0, qpos_noise=0.01, qvel_noise=0.3, ctrl_noise=1.5, exact_world0=False)
  sap2 = mjcf.load_string(util.CONVEX_XML)
  sap2.opt.broadphase, sap2.opt.broadphase_filter = 2, 1 | 8
  yield "convex_sap", sap2, dict(nconmax=64, njmax=256, key=None, qpos_noise=0.004, qvel_noise=0.05, ctrl_noise=0.0, exact_world0=False)
  rk = util.MIXED_XML.replace('<option timestep="0.004"', '<option integrator="RK4" timestep="0.004"')
  yield "mixed_rk4", mjcf.load_string(rk), dict(nconmax=32, njmax=128, key=0, qpos_noise=0.01, qvel_noise=0.3, ctrl_noise=1.5, exact_world0=False)
  yield "sensors", mjcf.load_string(util.sensor_xml()), dict(nconmax=32, njmax=128, key=0, qpos_noise=0.01, qvel_noise=0.3, ctrl_noise=1.5, exact_world0=False)
  for integ in ("Euler", "implicitfast", "RK4"):  # stateful actuators (na > 0): forward.py:135-218, 756-963, derivative.py:142-164
    yield "actuators" + ("" if integ == "Euler" else "_" + integ.lower()), mjcf.load_string(util.actuators_xml(integ)), dict(nconmax=16, njmax=64, key=0, qpos_noise=0.05, qvel_noise=0.5, ctrl_noise=1.2, exact_world0=False)
  for integ in ("Euler", "implicitfast"):  # fixed tendons: limits, spring / damper, friction loss, equalities, tendon transmissions
    yield "tendons" + ("" if integ == "Euler" else "_" + integ.lower()), mjcf.load_string(util.tendon_xml(integ)), dict(nconmax=16, njmax=64, key=0, qpos_noise=0.05, qvel_noise=0.5, ctrl_noise=1.0, exact_world0=False)
  # fully implicit integrator (forward.py:578-600, derivative.py:321-584 RNE velocity derivative, smooth.py:3376 sparse LU): the mixed scene
  # has free / ball / slide / hinge joints with several joints on one body, the actuator scene velocity-dependent actuator forces, the
  # tendon scene tendon damping
  imp = util.MIXED_XML.replace('<option timestep="0.004"', '<option integrator="implicit" timestep="0.004"')
  yield "mixed_implicit", mjcf.load_string(imp), dict(nconmax=32, njmax=128, key=0, qpos_noise=0.01, qvel_noise=1.0, ctrl_noise=1.5, exact_world0=False)
  yield "actuators_implicit", mjcf.load_string(util.actuators_xml("implicit")), dict(nconmax=16, njmax=64, key=0, qpos_noise=0.05, qvel_noise=0.5, ctrl_noise=1.2, exact_world0=False)
  yield "tendons_implicit", mjcf.load_string(util.tendon_xml("implicit")), dict(nconmax=16, njmax=64, key=0, qpos_noise=0.05, qvel_noise=0.5, ctrl_noise=1.0, exact_world0=False)
  yield "mesh", mjcf.load_string(util.mesh_xml()), dict(nconmax=64, njmax=256, key=None, qpos_noise=0.0004, qvel_noise=0.05, ctrl_noise=0.0, exact_world0=False)
  yield "equality", mjcf.load_string(util.EQUALITY_XML), dict(nconmax=16, njmax=64, key=0, qpos_noise=0.02, qvel_noise=0.5, ctrl_noise=0.5, exact_world0=False)
  three = mjcf.load_any(util.THREE_HUMANOIDS)
  three.opt.jacobian = 1  # sparse: the reference does not run dense above nv = 60; snapshot() stores efc_J densified
  yield "three_humanoids", three, dict(nconmax=100, njmax=192, key=0, qpos_noise=0.003, qvel_noise=0.05, ctrl_noise=0.3, exact_world0=False)
  yield "g1", mjcf.load_any(util.G1), dict(nconmax=48, njmax=192, key=0, qpos_noise=0.02, qvel_noise=0.2, ctrl_noise=0.3)


def snapshot(mjm, d, out, tag):
  for f in FIELDS:
    a = getattr(d, f, None)
    if a is not None and a.a is not None:
      out[f"{tag}/{f}"] = a.numpy()
  for f in EFC:
    a = getattr(d.efc, f)
    if a is not None and a.a is not None:
      out[f"{tag}/efc_{f}"] = a.numpy()
  if getattr(d.efc, "J_rownnz", None) is not None and d.efc.J_rownnz.a is not None and d.efc.J_rownnz.a.size and d.efc.J.numpy().shape[1] == 1:
    # CSR Jacobian (types.py:2021-2072) -> dense (nworld, njmax, nv) so that fixtures have one layout
    J, nnz, adr, col = d.efc.J.numpy(), d.efc.J_rownnz.numpy(), d.efc.J_rowadr.numpy(), d.efc.J_colind.numpy()
    nworld, njmax = nnz.shape
    dense = np.zeros((nworld, njmax, mjm.nv))
    nefc = d.nefc.numpy()
    for w in range(nworld):
      for r in range(int(nefc[w])):
        for k in range(int(nnz[w, r])):
          dense[w, r, col[w, 0, adr[w, r] + k]] = J[w, 0, adr[w, r] + k]
    out[f"{tag}/efc_J"] = dense
  nacon = int(d.nacon.numpy()[0])
  out[f"{tag}/nacon"] = np.array(nacon)
  for f in CON:
    out[f"{tag}/con_{f}"] = getattr(d.contact, f).numpy()[:nacon]
  for f in ("qpos", "qvel", "qacc_warmstart", "time", "act"):
    a = getattr(d, f, None)
    if a is not None and a.a is not None:
      out[f"{tag}/{f}"] = a.numpy()


def main(only=None):
  wp, ref = ref_runner.setup()
  io, fwd = ref["io"], ref["forward"]
  for name, mjm, cfg in scenes():
    if only and name not in only:
      continue
    t0 = time.time()
    nconmax, njmax = cfg.pop("nconmax"), cfg.pop("njmax")
    key = cfg.pop("key")
    NWORLD = cfg.pop("nworld", 3)
    qpos, qvel, ctrl, warm = util.seeded_state(mjm, NWORLD, key=key, seed=1234, **cfg)
    if name == "humanoid_trunc":  # lying on its back, torso 4.6 cm above the floor (no geom exactly tangent to it): head, torso, arms and legs all touch
      qpos[:, :3] = [0.0, 0.0, 0.046]
      qpos[:, 3:7] = [np.cos(np.pi / 4), 0.0, np.sin(np.pi / 4), 0.0]
    elif name.startswith("humanoid"):  # keep the feet on the floor: exact root pose, pushed 0.5 mm deeper per world
      qpos[:, :7] = mjm.key_qpos[0][:7]
      qpos[:, 2] -= 0.0005 * np.arange(NWORLD)
    if name == "three_humanoids":  # all three in the squat pose (keys 0, 3, 6 each pose one of them), roots exact, feet on the floor
      for i in range(3):
        sl = slice(28 * i, 28 * (i + 1))
        qpos[:, sl] += mjm.key_qpos[3 * i][sl] - mjm.key_qpos[0][sl]
        qpos[:, 28 * i : 28 * i + 7] = mjm.key_qpos[3 * i][28 * i : 28 * i + 7]
        qpos[:, 28 * i + 2] -= 0.0005 * np.arange(NWORLD)
    f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)  # inputs exactly representable in fp32
    qpos, qvel, ctrl, warm = f32(qpos), f32(qvel), f32(ctrl), f32(warm)
    ad = ref_runner.MjModelAdapter(mjm)
    m = io.put_model(ad)
    if getattr(mjm.opt, "broadphase", 0):  # put_model picks the broadphase from the pair count (io.py:631-636); override like a user would
      m.opt.broadphase = int(mjm.opt.broadphase)
    if hasattr(mjm.opt, "broadphase_filter"):
      m.opt.broadphase_filter = int(mjm.opt.broadphase_filter)
    d = io.make_data(ad, nworld=NWORLD, nconmax=nconmax, njmax=njmax)
    d.qpos.a[...] = qpos; d.qvel.a[...] = qvel; d.qacc_warmstart.a[...] = warm
    if mjm.nu:
      d.ctrl.a[...] = ctrl
    extra = {}
    if getattr(mjm, "na", 0):
      act = util.seeded_act(mjm, NWORLD)
      d.act.a[...] = act
      extra["in/act"] = act
    if getattr(mjm, "nmocap", 0):  # move the mocap bodies away from their model pose, differently per world
      rng = np.random.default_rng(77)
      mp = d.mocap_pos.numpy() + f32(0.03 * rng.uniform(-1, 1, (NWORLD, mjm.nmocap, 3)))
      mq = d.mocap_quat.numpy() + f32(0.1 * rng.uniform(-1, 1, (NWORLD, mjm.nmocap, 4)))
      mq = f32(mq / np.linalg.norm(mq, axis=-1, keepdims=True))
      d.mocap_pos.a[...] = mp; d.mocap_quat.a[...] = mq
      extra.update({"in/mocap_pos": mp, "in/mocap_quat": mq})
    out = {**extra, "in/qpos": qpos, "in/qvel": qvel, "in/ctrl": ctrl, "in/qacc_warmstart": warm, "in/nconmax": np.array(nconmax), "in/njmax": np.array(njmax)}
    fwd.forward(m, d)
    snapshot(mjm, d, out, "forward")
    out["forward/overflow"] = d.overflow.numpy() if getattr(d, "overflow", None) is not None else np.zeros(NWORLD, dtype=np.int32)
    for s in range(NSTEP):
      fwd.step(m, d)
      snapshot(mjm, d, out, f"step{s}")
      out[f"step{s}/overflow"] = d.overflow.numpy()
    path = os.path.join(ROOT, "tests", "golden", f"pipeline_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: nefc {out['forward/nefc'].ravel()}, nacon {int(out['forward/nacon'])}, niter {out['forward/solver_niter'].ravel()}, "
          f"{os.path.getsize(path) // 1024} KiB, {time.time() - t0:.1f} s; MjModel fallbacks: {len(ad.missing)}")


if __name__ == "__main__":
  # One process per scene: the reference keeps process-global kernel state (collision_primitive.py:1516 accumulates the
  # primitive pair types of every model the process has seen, so a box-box primitive from one model would also run for the next).
  names = sys.argv[1:] or [n for n, _, _ in scenes()]
  if len(names) == 1:
    main(names)
  else:
    import subprocess

    for n in names:
      subprocess.check_call([sys.executable, os.path.abspath(__file__), n])
