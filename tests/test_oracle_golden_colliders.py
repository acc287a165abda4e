"""Oracle narrowphase vs golden vectors produced by executing the reference's own collision_primitive_core.py
(tools/make_reference_goldens.py, run where /root/reference exists).  Every primitive pair function, 48 poses each."""

import json
import os

import numpy as np
import pytest

from tests import util

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_colliders.json")


def scene_xml(pair, gold):
  t1, s1, t2, s2 = pair["type1"], pair["size1"], pair["type2"], pair["size2"]
  half = gold["margin"] / 2
  sz = lambda s: " ".join(repr(float(x)) for x in s)
  if t1 == "plane":
    first = f'<geom name="g1" type="plane" size="{sz(s1)}" pos="{sz(gold["plane_pos"])}" quat="{sz(gold["plane_quat"])}" margin="{half}"/>'
  else:
    first = f'<body name="a"><freejoint/><geom name="g1" type="{t1}" size="{sz(s1)}" margin="{half}"/></body>'
  return f"""
<mujoco>
  <option><flag nativeccd="disable"/></option>
  <worldbody>
    {first}
    <body name="b"><freejoint/><geom name="g2" type="{t2}" size="{sz(s2)}" margin="{half}"/></body>
  </worldbody>
</mujoco>"""


@pytest.fixture(scope="module")
def gold():
  with open(GOLD) as f:
    return json.load(f)


def load_cases(gold, name):
  from mujoco_warp_b200._src import mjcf

  pair = gold["pairs"][name]
  mjm = mjcf.load_string(scene_xml(pair, gold))
  cases = pair["cases"]
  qpos = []
  for c in cases:
    q = [] if pair["type1"] == "plane" else list(c["pos1"]) + list(c["quat1"])
    qpos.append(q + list(c["pos2"]) + list(c["quat2"]))
  return mjm, cases, np.array(qpos)


PAIR_NAMES = ["plane_sphere", "plane_capsule", "plane_ellipsoid", "plane_cylinder", "plane_box", "sphere_sphere", "sphere_capsule",
              "sphere_cylinder", "sphere_box", "capsule_capsule", "capsule_box", "box_box"]


@pytest.mark.parametrize("name", PAIR_NAMES)
def test_oracle_matches_reference_collider(built, gold, name):
  mjm, cases, qpos = load_cases(gold, name)
  o = util.make_oracle(mjm, len(cases), 8, 64)
  o.set_state(qpos=qpos)
  o.forward()
  d = o.d
  assert (d["overflow"] == 0).all()
  margin = gold["margin"]
  total = 0
  for w, c in enumerate(cases):
    keep = [i for i, x in enumerate(c["dist"]) if x < margin]  # write_contact: dist < margin + gap (collision_primitive.py:198)
    n = d["ncon"][w]
    assert n == len(keep), f"{name} case {w}: {n} contacts, reference {len(keep)}"
    np.testing.assert_array_equal(d["con_geomcollisionid"][w, :n], keep)
    for k, i in enumerate(keep):
      np.testing.assert_allclose(d["con_dist"][w, k], c["dist"][i], atol=1e-9, err_msg=f"{name} case {w} dist[{i}]")
      np.testing.assert_allclose(d["con_pos"][w, k], c["pos"][i], atol=1e-9, err_msg=f"{name} case {w} pos[{i}]")
      np.testing.assert_allclose(d["con_frame"][w, k], c["frame"][i], atol=1e-8, err_msg=f"{name} case {w} frame[{i}]")
    total += n
  assert total >= 15, "golden set must exercise the collider"
