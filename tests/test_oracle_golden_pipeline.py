"""fp64 oracle vs trajectories produced by the reference's own pipeline (tools/make_pipeline_goldens.py: unmodified
reference sources executed on the CPU through tools/warp_shim.py).  forward() fields, constraint rows, solver output and
NSTEP steps of state, per scene."""

import glob
import os

import numpy as np
import pytest

from tests import util

GOLD_DIR = os.path.join(os.path.dirname(__file__), "golden")
SCENES = sorted(os.path.basename(p)[len("pipeline_"):-4] for p in glob.glob(os.path.join(GOLD_DIR, "pipeline_*.npz")))

SMOOTH = ["ten_length", "ten_velocity", "xpos", "xquat", "xmat", "xipos", "ximat", "xanchor", "xaxis", "geom_xpos", "geom_xmat", "site_xpos", "site_xmat", "cam_xpos", "cam_xmat",
          "light_xpos", "light_xdir", "subtree_com", "cdof", "cinert", "crb", "M", "actuator_length", "actuator_moment", "actuator_velocity", "cvel",
          "cdof_dot", "qfrc_bias", "qfrc_spring", "qfrc_damper", "qfrc_gravcomp", "qfrc_passive", "actuator_force", "qfrc_actuator", "qfrc_smooth", "qacc_smooth"]


def load_scene(name):
  from mujoco_warp_b200._src import constants as C
  from mujoco_warp_b200._src import mjcf
  from tests.test_gpu_colliders import BOX_XML

  if name.startswith("humanoid"):
    mjm = mjcf.load_any(util.HUMANOID)
    if name.endswith("elliptic"):
      mjm.opt.cone = C.CONE_ELLIPTIC
    if name.endswith("cg"):
      mjm.opt.solver = C.SOL_CG
  elif name == "mixed_elliptic_cg":
    mjm = mjcf.load_string(util.MIXED_XML.replace('<option timestep="0.004"', '<option solver="CG" cone="elliptic" timestep="0.004"'))
  elif name == "mixed":
    mjm = mjcf.load_string(util.MIXED_XML)
  elif name == "mixed_implicit":
    mjm = mjcf.load_string(util.MIXED_XML.replace('<option timestep="0.004"', '<option integrator="implicit" timestep="0.004"'))
  elif name == "mixed_elliptic":
    mjm = mjcf.load_string(util.MIXED_XML.replace('<option timestep="0.004"', '<option cone="elliptic" impratio="2" timestep="0.004"'))
  elif name == "boxes":
    mjm = mjcf.load_string(BOX_XML)
  elif name in ("boxccd", "boxccd_mixed"):
    mjm = mjcf.load_string(util.boxccd_xml(name.endswith("mixed")))
  elif name == "mesh":
    mjm = mjcf.load_string(util.mesh_xml())
  elif name == "sensors":
    mjm = mjcf.load_string(util.sensor_xml())
  elif name.startswith("tendons"):
    mjm = mjcf.load_string(util.tendon_xml("implicitfast" if name.endswith("implicitfast") else "implicit" if name.endswith("implicit") else "Euler"))
  elif name.startswith("actuators"):
    mjm = mjcf.load_string(util.actuators_xml({"actuators": "Euler", "actuators_implicitfast": "implicitfast", "actuators_rk4": "RK4", "actuators_implicit": "implicit"}[name]))
  elif name == "mixed_rk4":
    mjm = mjcf.load_string(util.MIXED_XML.replace('<option timestep="0.004"', '<option integrator="RK4" timestep="0.004"'))
  elif name == "mixed_sap":
    mjm = mjcf.load_string(util.MIXED_XML)
    mjm.opt.broadphase = 1
  elif name == "convex_sap":
    mjm = mjcf.load_string(util.CONVEX_XML)
    mjm.opt.broadphase, mjm.opt.broadphase_filter = 2, 1 | 8
  elif name == "convex":
    mjm = mjcf.load_string(util.CONVEX_XML)
  elif name == "pairs":
    mjm = mjcf.load_string(util.pairs_xml())
  elif name == "passive":
    mjm = mjcf.load_string(util.passive_xml())
  elif name == "equality":
    mjm = mjcf.load_string(util.EQUALITY_XML)
  elif name == "three_humanoids":
    mjm = mjcf.load_any(util.THREE_HUMANOIDS)
  elif name == "g1":
    mjm = mjcf.load_any(util.G1)
  else:
    raise KeyError(name)
  return mjm


def close(name, got, want, tol):
  got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
  want = want.reshape(got.shape)
  scale = max(1.0, float(np.abs(want).max(initial=0.0)))
  err = float(np.abs(got - want).max(initial=0.0))
  assert err <= tol * scale, f"{name}: max |diff| {err:.3e} (scale {scale:.3e})"


def compare(tag, g, od, mjm, nworld, tol, solver_tol=1e-7, exact_iterations=True):
  nv = mjm.nv
  for f in SMOOTH:
    k = f"{tag}/{f}"
    if k in g and f in od and g[k].size:
      want = g[k]
      if f == "M":
        want = want[:, : mjm.nC] if want.ndim == 2 else want
      if f == "actuator_moment":
        want = want.reshape(nworld, -1)[:, : od[f].reshape(nworld, -1).shape[1]]
      close(k, od[f].reshape(nworld, -1), want.reshape(nworld, -1), tol)
  for f in ("ne", "nf", "nl", "nefc"):
    np.testing.assert_array_equal(od[f].reshape(-1), g[f"{tag}/{f}"].reshape(-1), err_msg=f"{tag}/{f}")
  wid = g[f"{tag}/con_worldid"]
  for w in range(nworld):
    ids = np.nonzero(wid == w)[0]
    n = int(od["ncon"][w])
    assert n == len(ids), f"{tag} world {w}: {n} contacts, reference {len(ids)}"
    np.testing.assert_array_equal(od["con_geom"][w, :n], g[f"{tag}/con_geom"][ids])
    np.testing.assert_array_equal(od["con_dim"][w, :n], g[f"{tag}/con_dim"][ids])
    np.testing.assert_array_equal(od["con_geomcollisionid"][w, :n], g[f"{tag}/con_geomcollisionid"][ids])
    for f in ("dist", "pos", "frame", "includemargin", "friction", "solref", "solreffriction", "solimp"):
      close(f"{tag}/con_{f}[w{w}]", od["con_" + f][w, :n], g[f"{tag}/con_{f}"][ids], tol)
    ne = int(od["nefc"][w])
    np.testing.assert_array_equal(od["efc_type"][w, :ne], g[f"{tag}/efc_type"][w, :ne])
    eid = g[f"{tag}/efc_id"][w, :ne].copy()
    is_con = od["efc_type"][w, :ne] >= 5
    if n:  # the reference's contact ids index the global pool (a world's contacts need not be contiguous there)
      local = {int(p): k for k, p in enumerate(ids)}
      eid[is_con] = [local[int(p)] for p in eid[is_con]]
    np.testing.assert_array_equal(od["efc_id"][w, :ne], eid)
    np.testing.assert_array_equal(od["con_efc_address"][w, :n, : g[f"{tag}/con_efc_address"].shape[1]], g[f"{tag}/con_efc_address"][ids][:, : od["con_efc_address"].shape[2]])
    close(f"{tag}/efc_J[w{w}]", od["efc_J"][w, :ne], g[f"{tag}/efc_J"][w, :ne, :nv], tol)
    for f in ("pos", "margin", "D", "vel", "aref", "frictionloss"):
      close(f"{tag}/efc_{f}[w{w}]", od["efc_" + f][w, :ne], g[f"{tag}/efc_{f}"][w, :ne], tol)
    # solver: both sides iterate the same algorithm in double precision
    close(f"{tag}/efc_force[w{w}]", od["efc_force"][w, :ne], g[f"{tag}/efc_force"][w, :ne], solver_tol)
    if exact_iterations:
      np.testing.assert_array_equal(od["efc_state"][w, :ne], g[f"{tag}/efc_state"][w, :ne])
  if f"{tag}/sensordata" in g and g[f"{tag}/sensordata"].size:  # position / velocity sensors to tol, acceleration-stage ones follow qacc
    close(f"{tag}/sensordata", od["sensordata"], g[f"{tag}/sensordata"], solver_tol)
    for f in ("subtree_linvel", "subtree_angmom"):
      close(f"{tag}/{f}", od[f].reshape(nworld, -1), g[f"{tag}/{f}"].reshape(nworld, -1), tol)
    if f"{tag}/cfrc_ext" in g and np.abs(g[f"{tag}/cfrc_ext"]).max() > 0:  # rne_postconstraint ran (force / torque / accelerometer sensors)
      close(f"{tag}/cfrc_ext", od["cfrc_ext"].reshape(nworld, -1), g[f"{tag}/cfrc_ext"].reshape(nworld, -1), solver_tol * 10)
  close(f"{tag}/qacc", od["qacc"], g[f"{tag}/qacc"], solver_tol)
  close(f"{tag}/qfrc_constraint", od["qfrc_constraint"], g[f"{tag}/qfrc_constraint"], solver_tol)
  if exact_iterations:
    # the answers above already agree to 1e-7; a termination test that lands within rounding of the tolerance may add one
    # (idempotent) iteration on either side
    dn = np.abs(od["solver_niter"].reshape(-1) - g[f"{tag}/solver_niter"].reshape(-1))
    assert dn.max() <= 1 and (dn > 0).sum() <= 1, (od["solver_niter"].reshape(-1), g[f"{tag}/solver_niter"].reshape(-1))


@pytest.mark.parametrize("name", SCENES)
def test_oracle_matches_reference_pipeline(built, name):
  g = np.load(os.path.join(GOLD_DIR, f"pipeline_{name}.npz"))
  mjm = load_scene(name)
  nworld = g["in/qpos"].shape[0]
  o = util.make_oracle(mjm, nworld, int(g["in/nconmax"]), int(g["in/njmax"]))
  o.set_state(qpos=g["in/qpos"], qvel=g["in/qvel"], qacc_warmstart=g["in/qacc_warmstart"])
  if mjm.nu:
    o.set_state(ctrl=g["in/ctrl"])
  if "in/mocap_pos" in g:
    o.d["mocap_pos"][:] = g["in/mocap_pos"]; o.d["mocap_quat"][:] = g["in/mocap_quat"]
  if "in/act" in g:
    o.set_state(act=g["in/act"])
  o.forward()
  # capacity-overflow scenes: the oracle (like the CUDA collision kernel) raises the bit where the truncation happens, the reference
  # at the end of step() (forward.py:247-271); the bits are sticky, so both agree after a step (checked below)
  trunc = "step0/overflow" in g and int(g["step0/overflow"].max()) != 0
  assert trunc or (o.d["overflow"] == 0).all()
  # CG takes tens of iterations: rounding differences grow along the conjugate directions, so iteration counts can differ by
  # a few and the two (equally converged) answers agree to the solver tolerance rather than to rounding
  cg = name.endswith("cg")
  # mesh scene: cubes resting flat on cubes give four redundant contacts per face, whose force split is not unique -- the two sides agree
  # on it to the solver tolerance, not to rounding (geometry, Jacobians and every other constraint field still match to 1e-9)
  loose = cg or name == "mesh"
  compare("forward", g, o.d, mjm, nworld, 1e-9, solver_tol=5e-3 if cg else (2e-4 if loose else 1e-7), exact_iterations=not loose)
  if "in/act" in g:  # stateful actuators: activation derivatives of forward(), activations after every step
    close("forward/act_dot", o.d["act_dot"], g["forward/act_dot"], 1e-9)
  if cg:
    assert (np.abs(o.d["solver_niter"].reshape(-1) - g["forward/solver_niter"].reshape(-1)) <= 5).all()
  s = 0
  while f"step{s}/qpos" in g:
    o.step()
    close(f"step{s}/qpos", o.d["qpos"], g[f"step{s}/qpos"], 1e-5 if cg else 1e-6)
    # after the first step the two sides' inputs differ at rounding level, which can flip a borderline termination test
    # (one Newton iteration more or less); both results are converged to the solver tolerance (1e-6, scaled)
    close(f"step{s}/qvel", o.d["qvel"], g[f"step{s}/qvel"], 2e-3 if cg else 1e-4)
    close(f"step{s}/qacc_warmstart", o.d["qacc_warmstart"], g[f"step{s}/qacc_warmstart"], 5e-3 if cg else 2e-4)
    close(f"step{s}/time", o.d["time"], g[f"step{s}/time"], 1e-12)
    if "in/act" in g:
      close(f"step{s}/act", o.d["act"], g[f"step{s}/act"], 1e-7)
    np.testing.assert_array_equal(o.d["nefc"].reshape(-1), g[f"step{s}/nefc"].reshape(-1), err_msg=f"step{s}/nefc")
    if trunc:
      np.testing.assert_array_equal(o.d["overflow"].reshape(-1) & ~(1 << 10), g[f"step{s}/overflow"].reshape(-1) & ~(1 << 10), err_msg=f"step{s}/overflow")
    s += 1
  assert s >= 3
