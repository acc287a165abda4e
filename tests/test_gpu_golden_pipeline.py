"""CUDA path (through the C-ABI) vs the trajectories produced by the reference's own pipeline
(tests/golden/pipeline_*.npz, see tools/make_pipeline_goldens.py).  fp32 kernels against the reference evaluated in double:
smooth / constraint quantities within 5e-4 (the reference's own test tolerance, smooth_test.py:32), solver outputs within
5e-3 of the force scale, integer outputs exact."""

import os

import numpy as np
import pytest
import torch

from tests import util
from tests.test_oracle_golden_pipeline import GOLD_DIR, SCENES, SMOOTH, load_scene

pytestmark = pytest.mark.gpu


def close(name, got, want, atol, rtol=0.0):
  got = np.asarray(got, dtype=np.float64)
  want = np.asarray(want, dtype=np.float64).reshape(got.shape)
  util.assert_close(name, got, want, atol=atol, rtol=rtol)


GPU_SCENES = list(SCENES)  # incl. `mesh`: hull-vertex support function, hill climbing, mesh multi-contact, plane-mesh (k_collision_mesh.cu)


@pytest.mark.parametrize("name", GPU_SCENES)
def test_gpu_matches_reference_pipeline(built, name):
  import mujoco_warp_b200 as mjw

  g = np.load(os.path.join(GOLD_DIR, f"pipeline_{name}.npz"))
  mjm = load_scene(name)
  m = mjw.put_model(mjm)
  nworld = g["in/qpos"].shape[0]
  d = mjw.make_data(mjm, nworld=nworld, nconmax=int(g["in/nconmax"]), njmax=int(g["in/njmax"]), m=m)
  f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
  d.qpos.copy_(f32(g["in/qpos"])); d.qvel.copy_(f32(g["in/qvel"])); d.qacc_warmstart.copy_(f32(g["in/qacc_warmstart"]))
  if mjm.nu:
    d.ctrl.copy_(f32(g["in/ctrl"]))
  if "in/mocap_pos" in g:
    d.mocap_pos.copy_(f32(g["in/mocap_pos"])); d.mocap_quat.copy_(f32(g["in/mocap_quat"]))
  if "in/act" in g:  # stateful actuators
    d.act.copy_(f32(g["in/act"]))
  mjw.forward(m, d)
  torch.cuda.synchronize()
  trunc = "step0/overflow" in g and int(g["step0/overflow"].max()) != 0  # capacity-overflow scene: bits are raised where the truncation happens
  assert trunc or (d.overflow.cpu().numpy() == 0).all()
  nv, tag = mjm.nv, "forward"
  # flat-on-flat convex contacts (cylinder cap on a box face, crossed cylinders ...) have a whole patch of valid witness points:
  # EPA in fp32 and in double stop at different ones, so positions (and everything downstream of the torque arm) get a looser band
  flat = name.startswith("convex") or name.startswith("boxccd") or name.startswith("mesh")
  ptol = 5e-3 if flat else 5e-4
  for f in SMOOTH:
    k = f"{tag}/{f}"
    if k not in g or not g[k].size or not hasattr(d, f):
      continue
    got = getattr(d, f).cpu().numpy().reshape(nworld, -1)
    want = g[k]
    if f == "M":
      want = want[:, : mjm.nC]
    want = want.reshape(nworld, -1)[:, : got.shape[1]]
    scale = max(1.0, float(np.abs(want).max()))
    close(k, got, want, atol=5e-4 * scale)
  for f in ("ne", "nf", "nl", "nefc"):
    np.testing.assert_array_equal(getattr(d, f).cpu().numpy().reshape(-1), g[f"{tag}/{f}"].reshape(-1), err_msg=f)
  assert int(d.nacon.cpu()[0]) == int(g[f"{tag}/nacon"])
  wid = g[f"{tag}/con_worldid"]
  J = util.dense_J(d)
  fscale = max(1.0, float(np.abs(g[f"{tag}/efc_force"]).max()))
  for w in range(nworld):
    ids = util.world_contacts(d, w)
    ref_ids = np.nonzero(wid == w)[0]
    assert len(ids) == len(ref_ids)
    c = d.contact
    ne = min(int(g[f"{tag}/nefc"].reshape(-1)[w]), int(g["in/njmax"]))  # rows beyond njmax are dropped (constraint.py:2048,2712)
    rows = np.arange(ne)  # got row that corresponds to reference row i
    if name.startswith("mesh") and len(ids):
      # a mesh lying flat on the floor has several hull vertices within rounding of the deepest one: plane_convex then picks the same
      # set of (up to four) vertices in fp32 and in double, but may start from a different one.  Contacts of one geom pair are
      # therefore matched by position, and their constraint rows are permuted accordingly, before the field-by-field comparison.
      gg, wg = c.geom[ids].cpu().numpy(), g[f"{tag}/con_geom"][ref_ids]
      gp, wp = c.pos[ids].cpu().numpy().astype(np.float64), g[f"{tag}/con_pos"][ref_ids]
      perm, used = [], set()
      for i in range(len(ref_ids)):
        cand = [j for j in range(len(ids)) if j not in used and (gg[j] == wg[i]).all()]
        assert cand, f"world {w}: no contact left for geom pair {wg[i]}"
        j = min(cand, key=lambda j: float(np.abs(gp[j] - wp[i]).sum()))
        used.add(j); perm.append(j)
      perm = np.asarray(perm)
      gadr = c.efc_address[ids].cpu().numpy()[perm]
      wadr = g[f"{tag}/con_efc_address"][ref_ids]
      for a_got, a_want in zip(gadr, wadr):
        for r_got, r_want in zip(a_got, a_want):
          assert (r_got >= 0) == (r_want >= 0)
          if r_want >= 0:
            rows[r_want] = r_got
      ids = ids[torch.as_tensor(perm, device=ids.device)] if isinstance(ids, torch.Tensor) else np.asarray(ids)[perm]
    else:
      np.testing.assert_array_equal(c.geomcollisionid[ids].cpu().numpy(), g[f"{tag}/con_geomcollisionid"][ref_ids])
    np.testing.assert_array_equal(c.geom[ids].cpu().numpy(), g[f"{tag}/con_geom"][ref_ids])
    np.testing.assert_array_equal(c.dim[ids].cpu().numpy(), g[f"{tag}/con_dim"][ref_ids])
    for f in ("dist", "pos", "frame", "includemargin", "friction", "solref", "solimp") if len(ids) else ():
      tol = ptol if f in ("pos", "frame") else 5e-4
      close(f"con_{f}[w{w}]", getattr(c, f)[ids].cpu().numpy().reshape(len(ids), -1), g[f"{tag}/con_{f}"][ref_ids].reshape(len(ids), -1), atol=tol, rtol=5e-4)
    np.testing.assert_array_equal(d.efc.type[w, :ne].cpu().numpy()[rows], g[f"{tag}/efc_type"][w, :ne])
    close(f"efc_J[w{w}]", J[w, :ne, :nv][rows], g[f"{tag}/efc_J"][w, :ne, :nv], atol=ptol, rtol=5e-4)
    for f in ("pos", "margin", "vel", "frictionloss"):
      close(f"efc_{f}[w{w}]", getattr(d.efc, f)[w, :ne].cpu().numpy()[rows], g[f"{tag}/efc_{f}"][w, :ne], atol=5e-4, rtol=5e-4)
    close(f"efc_D[w{w}]", d.efc.D[w, :ne].cpu().numpy()[rows], g[f"{tag}/efc_D"][w, :ne], atol=1e-3, rtol=2e-3)
    # aref = -k imp pos - b vel with k ~ 1e4: an fp32 penetration depth (error ~2e-6) moves aref by ~1e-2
    close(f"efc_aref[w{w}]", d.efc.aref[w, :ne].cpu().numpy()[rows], g[f"{tag}/efc_aref"][w, :ne], atol=2e-3, rtol=1e-2)
    close(f"efc_force[w{w}]", d.efc.force[w, :ne].cpu().numpy()[rows], g[f"{tag}/efc_force"][w, :ne], atol=(5e-2 if flat else 5e-3) * fscale)
  scale = max(1.0, float(np.abs(g[f"{tag}/qacc"]).max()))
  close("qacc", d.qacc.cpu().numpy(), g[f"{tag}/qacc"], atol=(5e-2 if flat else 5e-3) * scale)
  if f"{tag}/sensordata" in g and g[f"{tag}/sensordata"].size:
    # position / velocity sensors at the smooth-field tolerance; accelerometers inherit the solver's qacc band times the lever arm
    stage = np.repeat(np.asarray(mjm.sensor_needstage), np.asarray(mjm.sensor_dim))
    got, want = d.sensordata.cpu().numpy(), g[f"{tag}/sensordata"]
    close("sensordata[pos, vel]", got[:, stage < 3], want[:, stage < 3], atol=5e-4, rtol=5e-4)
    close("sensordata[acc]", got[:, stage == 3], want[:, stage == 3], atol=5e-3 * scale, rtol=5e-3)
    for f in ("subtree_linvel", "subtree_angmom"):
      close(f, getattr(d, f).cpu().numpy().reshape(nworld, -1), g[f"{tag}/{f}"].reshape(nworld, -1), atol=5e-4, rtol=5e-4)
    if f"{tag}/cfrc_ext" in g and np.abs(g[f"{tag}/cfrc_ext"]).max() > 0:
      fs = max(1.0, float(np.abs(g[f"{tag}/cfrc_ext"]).max()))
      close("cfrc_ext", d.cfrc_ext.cpu().numpy().reshape(nworld, -1), g[f"{tag}/cfrc_ext"].reshape(nworld, -1), atol=5e-3 * fs, rtol=5e-3)
  # Stepped states, teacher-forced: step s starts from the REFERENCE's state after step s-1 (qpos, qvel, warm start, time), so a
  # rounding-level difference cannot grow chaotically over the steps and the band is one step's worth of the qacc band above:
  # |dqvel| <= dt * |dqacc|, |dqpos| <= dt * |dqvel|.  A world is compared when its row count equals the reference's for that step.
  # The box scenes contain a knife-edge multi-contact decision: the pair (4, 5) of boxccd -- a box turned 45 degrees under another,
  # tilted by 8e-4 rad -- alternates between a face patch of 4 contacts and an edge of 2 from step to step in the reference itself
  # (112 / 104 / 112 rows), and which step takes which side depends on the last bits of the pose (the GPU's xmat differs from the
  # fp32 oracle's in the 8th digit; the device routine fed the oracle's pose on the host reproduces the oracle's count).  Up to a
  # quarter of the world-steps of a flat-contact scene may differ that way; every other scene must match every step.
  dt = float(np.asarray(mjm.opt.timestep))
  if "in/act" in g:
    close("forward/act_dot", d.act_dot.cpu().numpy(), g["forward/act_dot"], atol=1e-4, rtol=1e-4)
  s, skipped = 0, 0
  while f"step{s}/qpos" in g:
    if s > 0 and "in/act" in g:
      d.act.copy_(f32(g[f"step{s - 1}/act"]))
    if s > 0 and f"step{s - 1}/qacc_warmstart" in g:
      d.qpos.copy_(f32(g[f"step{s - 1}/qpos"])); d.qvel.copy_(f32(g[f"step{s - 1}/qvel"]))
      d.qacc_warmstart.copy_(f32(g[f"step{s - 1}/qacc_warmstart"]))
      d.time.copy_(f32(np.asarray(g[f"step{s - 1}/time"]).reshape(-1)))
    mjw.step(m, d)
    torch.cuda.synchronize()
    same = np.ones(nworld, dtype=bool)
    if f"step{s}/nefc" in g and not trunc:
      same = d.nefc.cpu().numpy().reshape(-1) == g[f"step{s}/nefc"].reshape(-1)
      skipped += int((~same).sum())
    ascale = max(1.0, float(np.abs(g[f"step{s}/qacc"]).max())) if f"step{s}/qacc" in g else scale
    vtol = dt * (1e-2 if flat else 5e-3) * ascale + 1e-4
    close(f"step{s}/qvel", d.qvel.cpu().numpy()[same], g[f"step{s}/qvel"][same], atol=vtol, rtol=1e-3)
    close(f"step{s}/qpos", d.qpos.cpu().numpy()[same], g[f"step{s}/qpos"][same], atol=dt * vtol + 2e-5, rtol=1e-5)
    if "in/act" in g:
      close(f"step{s}/act", d.act.cpu().numpy()[same], g[f"step{s}/act"][same], atol=1e-5, rtol=1e-5)
    if name.endswith("_implicit"):  # fully implicit integrator: the LU factors of M - dt (qDeriv_smooth + d RNE / d qvel), D-structure
      want = g[f"step{s}/qLU"]
      assert d.qLU.shape == want.shape and np.abs(want).max() > 0
      close(f"step{s}/qLU", d.qLU.cpu().numpy()[same], want[same], atol=2e-3 * max(1.0, float(np.abs(want).max())), rtol=2e-3)
    s += 1
  assert skipped <= (nworld * s // 4 if flat else 0), f"{skipped} world-steps with a row count different from the reference's"
  # the line-search budget flag (1 << 10) may be raised in fp32 when the bracketing stalls at rounding level; nothing else may
  # (CG scenes run close to their iteration cap -- 41..48 of 50 in double -- so fp32 may also raise the iteration flag, 1 << 9)
  allowed = (1 << 10) | ((1 << 9) if name.endswith("cg") else 0)
  if f"step{s - 1}/overflow" in g:  # capacity-overflow scene: the reference's own overflow bits must be raised (njmax truncation -> NEFC)
    want_ovf = g[f"step{s - 1}/overflow"].reshape(-1).astype(np.int64)
    np.testing.assert_array_equal(d.overflow.cpu().numpy().astype(np.int64) & ~allowed, want_ovf & ~allowed)
    allowed |= int(want_ovf.max())
  assert s >= 3 and ((d.overflow.cpu().numpy() & ~allowed) == 0).all()
