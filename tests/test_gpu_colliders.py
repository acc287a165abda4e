"""GPU narrowphase (through the C-ABI) vs the reference-generated collider vectors and vs the oracle on a box scene."""

import json

import numpy as np
import pytest
import torch

from tests import util
from tests.test_oracle_golden_colliders import GOLD, PAIR_NAMES, load_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
  with open(GOLD) as f:
    return json.load(f)


@pytest.mark.parametrize("name", PAIR_NAMES)
def test_gpu_matches_reference_collider(built, gold, name):
  import mujoco_warp_b200 as mjw

  mjm, cases, qpos = load_cases(gold, name)
  m = mjw.put_model(mjm)
  nworld = len(cases)
  d = mjw.make_data(mjm, nworld=nworld, nconmax=8, njmax=64, m=m)
  d.qpos.copy_(torch.from_numpy(qpos.astype(np.float32)))
  mjw.kinematics(m, d)
  mjw.collision(m, d)
  torch.cuda.synchronize()
  assert (d.overflow.cpu().numpy() == 0).all()
  margin = gold["margin"]
  dist, pos, frame, cid = (x.cpu().numpy() for x in (d.contact.dist, d.contact.pos, d.contact.frame, d.contact.geomcollisionid))
  mismatched, total = [], 0
  for w, c in enumerate(cases):
    if name in ("capsule_capsule", "box_box") and w % 8 == 7:
      # exactly aligned poses are ties decided by fp32 rounding: parallel capsule axes (`abs(det) >= MJ_MINVAL`,
      # collision_primitive_core.py:158) and equal face separations of two stacked boxes (:640-645 picks the reference face)
      continue
    ids = util.world_contacts(d, w)
    keep = [i for i, x in enumerate(c["dist"]) if x < margin]
    borderline = any(abs(x - margin) < 1e-5 for x in c["dist"])
    if len(ids) != len(keep) or (cid[ids] != keep).any():
      assert borderline or name in ("box_box", "capsule_box"), f"{name} case {w}: contact ids {cid[ids]} vs reference {keep}"
      mismatched.append(w)  # fp32 picked another (near-tied) feature in the box clipping code
      continue
    ok = True
    for k, i in enumerate(keep):
      ok &= abs(dist[ids[k]] - c["dist"][i]) < 5e-5 and np.abs(pos[ids[k]] - c["pos"][i]).max() < 5e-5 and np.abs(frame[ids[k]].reshape(3, 3) - c["frame"][i]).max() < 2e-3
    if not ok:
      assert name in ("box_box", "capsule_box"), f"{name} case {w}: contact values differ from the reference"
      mismatched.append(w)
    total += len(keep)
  assert len(mismatched) <= 2, f"{name}: cases {mismatched} differ from the reference"
  assert total >= 10


BOX_XML = """
<mujoco>
  <option timestep="0.002" iterations="50"><flag nativeccd="disable"/></option>
  <default><geom friction="0.9 0.01 0.002"/></default>
  <worldbody>
    <geom name="floor" type="plane" size="0 0 .05" condim="3" contype="7" conaffinity="7"/>
    <body pos="0 0 0.099"><freejoint/><geom type="box" size="0.2 0.2 0.1" density="500"/></body>
    <body pos="0.05 0.02 0.278"><freejoint/><geom type="box" size="0.1 0.1 0.08" density="500"/></body>
    <body pos="0.6 0 0.049" euler="90 0 0"><freejoint/><geom type="capsule" size="0.05 0.15"/></body>
    <body pos="0.6 0 0.178"><freejoint/><geom type="box" size="0.08 0.08 0.08" density="300"/></body>
    <body pos="-0.5 0.3 0.099"><freejoint/><geom type="cylinder" size="0.08 0.1" contype="2" conaffinity="2"/></body>
    <body pos="-0.5 0.3 0.268"><freejoint/><geom type="sphere" size="0.07" contype="3" conaffinity="3"/></body>
    <body pos="-0.5 -0.4 0.059"><freejoint/><geom type="ellipsoid" size="0.1 0.15 0.06" contype="4" conaffinity="4"/></body>
    <body pos="0.1 -0.6 0.198"><freejoint/><geom type="sphere" size="0.06"/></body>
    <body pos="0.1 -0.6 0.069"><freejoint/><geom type="box" size="0.1 0.1 0.07"/></body>
  </worldbody>
</mujoco>"""


def test_box_scene_forward_and_rollout(built):
  """(contype/conaffinity keep the cylinder and the ellipsoid away from box/capsule pairs, which the reference sends to GJK.)
  Boxes, a cylinder and an ellipsoid resting on the floor / on each other: contacts, efc rows and the solve match the oracle,
  then 40 steps stay in agreement."""
  import mujoco_warp_b200 as mjw

  mjm = mjw.mjcf.load_string(BOX_XML)
  m = mjw.put_model(mjm)
  nworld, nconmax, njmax = 8, 48, 200
  d = mjw.make_data(mjm, nworld=nworld, nconmax=nconmax, njmax=njmax, m=m)
  o = util.make_oracle(mjm, nworld, nconmax, njmax)
  rng = np.random.default_rng(3)
  qpos = np.tile(mjm.qpos0, (nworld, 1)).astype(np.float64)
  qvel = 0.05 * rng.uniform(-1, 1, (nworld, mjm.nv))
  for b in range(mjm.nbody - 1):
    qpos[1:, 7 * b : 7 * b + 2] += 0.005 * rng.uniform(-1, 1, (nworld - 1, 2))
    q = qpos[1:, 7 * b + 3 : 7 * b + 7] + 0.01 * rng.uniform(-1, 1, (nworld - 1, 4))
    qpos[1:, 7 * b + 3 : 7 * b + 7] = q / np.linalg.norm(q, axis=1, keepdims=True)
  f32 = lambda a: a.astype(np.float32)
  d.qpos.copy_(torch.from_numpy(f32(qpos))); d.qvel.copy_(torch.from_numpy(f32(qvel)))
  o.set_state(qpos=f32(qpos), qvel=f32(qvel))
  mjw.forward(m, d)
  o.forward()
  torch.cuda.synchronize()
  od = o.d
  assert (od["overflow"] == 0).all() and (d.overflow.cpu().numpy() == 0).all()
  assert od["ncon"].min() >= 15
  same = 0
  for w in range(1, nworld):  # world 0 sits exactly axis-aligned: clipping ties are decided by rounding
    ids = util.world_contacts(d, w)
    n = int(od["ncon"][w])
    if len(ids) != n or (d.contact.geom[ids].cpu().numpy() != od["con_geom"][w, :n]).any():
      continue
    same += 1
    util.assert_close(f"dist[w{w}]", d.contact.dist[ids].cpu().numpy(), od["con_dist"][w, :n], atol=2e-5, rtol=0)
    util.assert_close(f"pos[w{w}]", d.contact.pos[ids].cpu().numpy(), od["con_pos"][w, :n], atol=5e-5, rtol=0)
    util.assert_close(f"frame[w{w}]", d.contact.frame[ids].cpu().numpy().reshape(n, 3, 3), od["con_frame"][w, :n], atol=2e-3, rtol=0)
    assert int(d.nefc[w]) == int(od["nefc"][w])
    scale = max(1.0, float(np.abs(od["qacc"][w]).max()))
    util.assert_close(f"qacc[w{w}]", d.qacc[w].cpu().numpy(), od["qacc"][w], atol=2e-2 * scale, rtol=0)
  assert same >= nworld - 2, same
  for i in range(40):
    mjw.step(m, d)
    o.step()
  torch.cuda.synchronize()
  assert (d.overflow.cpu().numpy() == 0).all()
  util.assert_close("qpos@40", d.qpos.cpu().numpy()[1:], od["qpos"][1:], atol=3e-3, rtol=0)
  assert np.isfinite(d.qpos.cpu().numpy()).all()
