"""The reference's own known-answer tests for the broadphase (broadphase_test.py:44-342), transcribed as data: scenes, keyframes, filter
combinations and the expected number of candidate pairs are the reference's; they run through the oracle's collision stage for the
NXN rule and both sweep-and-prune variants.  (The CUDA path is held to the oracle's `ncollision` on the pipeline scenes.)"""
import numpy as np
import pytest

from mujoco_warp_b200._src import constants as C
from mujoco_warp_b200._src import mjcf
from tests import util

PLANE, SPHERE, AABB, OBB = 1, 2, 4, 8
FILTERS = [PLANE | SPHERE, PLANE | AABB, PLANE | OBB, PLANE | SPHERE | AABB, PLANE | SPHERE | OBB, PLANE | SPHERE | AABB | OBB]

FIVE_BODIES = """
<mujoco>
  <worldbody>
    <body><freejoint/><geom type="sphere" size="0.1"/></body>
    <body><freejoint/><geom type="sphere" size="0.1"/></body>
    <body><freejoint/><geom type="capsule" size="0.1 0.1"/></body>
    <body><freejoint/><geom type="sphere" size="0.1"/></body>
    <body>
      <freejoint/>
      <geom type="sphere" size="0.1"/>
      <geom type="sphere" size="0.1"/>
      <body><geom type="sphere" size="0.1"/><joint type="hinge"/></body>
    </body>
  </worldbody>
  <keyframe>
    <key qpos='0 0 0 1 0 0 0  1 0 0 1 0 0 0  2 0 0 1 0 0 0  3 0 0 1 0 0 0  4 0 0 1 0 0 0  0'/>
    <key qpos='0 0 0 1 0 0 0  .05 0 0 1 0 0 0  2 0 0 1 0 0 0  3 0 0 1 0 0 0  4 0 0 1 0 0 0  0'/>
    <key qpos='0 0 0 1 0 0 0  .01 0 0 1 0 0 0  .02 0 0 1 0 0 0  3 0 0 1 0 0 0  4 0 0 1 0 0 0  0'/>
    <key qpos='0 0 0 1 0 0 0  1 0 0 1 0 0 0  2 0 0 1 0 0 0  2 0 0 1 0 0 0  4 0 0 1 0 0 0  0'/>
  </keyframe>
</mujoco>"""

PLANE_CAPSULE_CAPSULE = """
<mujoco>
  <option gravity="0 0 0"/>
  <worldbody>
    <geom name="floor" size="10 10 .001" type="plane"/>
    <body>
      <geom type="capsule" size=".05 .1"/>
      <joint type="slide" axis="1 0 0"/><joint type="slide" axis="0 0 1"/><joint type="hinge" axis="0 1 0"/>
    </body>
    <body>
      <geom type="capsule" size=".05 .1"/>
      <joint type="slide" axis="1 0 0"/><joint type="slide" axis="0 0 1"/><joint type="hinge" axis="0 1 0"/>
    </body>
  </worldbody>
  <keyframe>
    <key qpos="-.5 .25 0 .5 .25 0"/>
    <key qpos="-.5 .075 1.57 .5 .25 0"/>
    <key qpos="-.075 .25 0 .075 .25 0"/>
    <key qpos="0 .25 .7853 0 .45 .7853"/>
  </keyframe>
</mujoco>"""


def run(mjm, qpos):
  """collision stage of the oracle on the given world states -> (ncollision per world, contact geom pairs per world)."""
  qpos = np.atleast_2d(np.asarray(qpos, dtype=np.float64))
  o = util.make_oracle(mjm, qpos.shape[0], 32, 64)
  o.set_state(qpos=qpos)
  o.forward()
  assert (o.d["overflow"] == 0).all()
  pairs = [sorted(map(tuple, o.d["con_geom"][w, : int(o.d["ncon"][w])].tolist())) for w in range(qpos.shape[0])]
  return np.asarray(o.d["ncollision"]).reshape(-1), pairs


def keyq(mjm, k):
  return np.asarray(mjm.key_qpos, dtype=np.float64).reshape(-1, mjm.nq)[k]


@pytest.mark.parametrize("filt", FILTERS)
@pytest.mark.parametrize("broadphase", [0, 1, 2])  # NXN, SAP_TILE, SAP_SEGMENTED
def test_broadphase_pair_counts(built, broadphase, filt):  # broadphase_test.py:52-176
  mjm = mjcf.load_string(FIVE_BODIES)
  mjm.opt.broadphase, mjm.opt.broadphase_filter = broadphase, filt
  n, _ = run(mjm, keyq(mjm, 0))
  assert n[0] == 0
  n, pairs = run(mjm, keyq(mjm, 1))
  assert n[0] == 1 and pairs[0] == [(0, 1)]
  n, pairs = run(mjm, keyq(mjm, 2))
  assert n[0] == 3 and pairs[0] == [(0, 1), (0, 2), (1, 2)]
  # two worlds, four candidate pairs: world 0 holds keyframe 1, world 1 keyframe 2
  n, pairs = run(mjm, np.stack([keyq(mjm, 1), keyq(mjm, 2)]))
  assert n.tolist() == [1, 3] and pairs == [[(0, 1)], [(0, 1), (0, 2), (1, 2)]]
  # geom type ordering: the sphere (geom 3) comes first in the pair with the capsule (geom 2)
  n, pairs = run(mjm, keyq(mjm, 3))
  assert n[0] == 1 and pairs[0] == [(3, 2)]
  # contype / conaffinity incompatibility
  mjm4 = mjcf.load_string(FIVE_BODIES)
  mjm4.opt.broadphase, mjm4.opt.broadphase_filter = broadphase, filt
  mjm4.geom_contype[:3] = 0
  n, _ = run(mjm4, keyq(mjm4, 1))
  assert n[0] == 0


@pytest.mark.parametrize("margin1,margin2,expected", [(0, 0, 0), (0, 0.011, 1), (0.011, 0, 1), (0.00999, 0, 0), (0, 0.00999, 0), (0.00999, 0.00999, 1)])
def test_broadphase_margin(built, margin1, margin2, expected):  # :178-207
  xml = f"""
<mujoco>
  <worldbody>
    <body><geom type="sphere" size=".1" margin="{margin1}"/><joint type="slide" axis="1 0 0"/></body>
    <body><geom type="sphere" size=".1" margin="{margin2}"/><joint type="slide" axis="1 0 0"/></body>
  </worldbody>
</mujoco>"""
  mjm = mjcf.load_string(xml)
  n, _ = run(mjm, [0.0, 0.21])
  assert n[0] == expected


@pytest.mark.parametrize("disable,expected", [(0, 0), (C.DSBL_FILTERPARENT, 1)])
def test_broadphase_filterparent(built, disable, expected):  # :209-233
  xml = """
<mujoco>
  <worldbody>
    <body>
      <geom type="sphere" size=".1"/><joint type="slide"/>
      <body><geom type="sphere" size=".1"/><joint type="slide"/></body>
    </body>
  </worldbody>
</mujoco>"""
  mjm = mjcf.load_string(xml)
  mjm.opt.disableflags = int(mjm.opt.disableflags) | disable
  n, _ = run(mjm, [0.0, 0.0])
  assert n[0] == expected


# (keyframe, filter, expected candidate pairs) -- broadphase_test.py:235-342
FILTER_CASES = [(0, PLANE | SPHERE, 0), (0, PLANE | AABB, 0), (0, PLANE | OBB, 0), (1, PLANE | SPHERE, 1), (1, PLANE, 2), (1, PLANE | OBB, 1),
                (2, PLANE | SPHERE, 1), (2, PLANE | AABB, 0), (2, PLANE | OBB, 0), (3, PLANE | SPHERE, 1), (3, PLANE | AABB, 1), (3, PLANE | OBB, 0)]


@pytest.mark.parametrize("key,filt,expected", FILTER_CASES)
def test_broadphase_filter(built, key, filt, expected):
  mjm = mjcf.load_string(PLANE_CAPSULE_CAPSULE)
  mjm.opt.broadphase_filter = filt
  n, _ = run(mjm, keyq(mjm, key))
  assert n[0] == expected
