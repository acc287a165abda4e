"""Known-answer tests of the reference's collision driver that need no MuJoCo (collision_driver_test.py:662-712 exclude table and plane /
tetrahedron, :713-880 explicit contact pairs, :949-975 minimum friction), transcribed as data and run against the oracle (and, for the
plane / mesh case, against the device routine compiled for the host)."""
import ctypes

import numpy as np
import pytest

from mujoco_warp_b200._src import io as mio
from mujoco_warp_b200._src import mjcf
from tests import util
from tests.test_device_ccd_mesh_on_host import hlib  # noqa: F401  (fixture: the CCD_MESH host build of the device header)

PAIR = '<pair geom1="geom1" geom2="geom2" margin="-1" gap="3" condim="6" friction="5 4 3 2 1" solref="-.25 -.5" solreffriction="2 4" solimp=".1 .2 .3 .4 .5"/>'
TWO = """
<mujoco>
  <worldbody>
    <body name="body1"><freejoint/><geom name="geom1" type="sphere" size=".1"{a}/></body>
    <body name="body2"><freejoint/><geom name="geom2" type="sphere" size=".1"{a}/></body>
  </worldbody>
  <contact>{c}</contact>
</mujoco>"""


def collide(mjm, qpos=None):
  o = util.make_oracle(mjm, 1, 16, 64)
  o.set_state(qpos=np.asarray(mjm.qpos0 if qpos is None else qpos, dtype=np.float64).reshape(1, -1))
  o.forward()
  return o.d


def test_no_pairs():  # :716-731
  mjm = mjcf.load_string('<mujoco><worldbody><body><freejoint/><geom type="sphere" size=".1"/></body></worldbody></mujoco>')
  t = mio.derive_tables(mjm)
  assert (np.asarray(t["nxn_pairid"]).reshape(-1, 2)[:, 0] == -1).all() if np.asarray(t["nxn_pairid"]).size else True


@pytest.mark.parametrize("attrs,extra", [("", ""), (' contype="0" conaffinity="0"', ""), ("", '<exclude body1="body1" body2="body2"/>')])
def test_contact_pair_overrides(built, attrs, extra):  # :732-880 plain pair / pair overrides contype+conaffinity / pair overrides exclude
  mjm = mjcf.load_string(TWO.format(a=attrs, c=extra + PAIR))
  t = mio.derive_tables(mjm)
  assert (np.asarray(t["nxn_pairid"]).reshape(-1, 2)[:, 0] == 0).all()
  d = collide(mjm)  # both spheres at the origin
  assert int(d["ncon"][0]) == 1
  assert d["con_includemargin"][0, 0] == -1
  assert d["con_dim"][0, 0] == 6
  np.testing.assert_allclose(d["con_friction"][0, 0], [5, 4, 3, 2, 1])
  np.testing.assert_allclose(d["con_solref"][0, 0], [-0.25, -0.5])
  np.testing.assert_allclose(d["con_solreffriction"][0, 0], [2.0, 4.0])
  np.testing.assert_allclose(d["con_solimp"][0, 0], [0.1, 0.2, 0.3, 0.4, 0.5])


def test_contact_exclude():  # :662-686
  xml = """
<mujoco>
  <worldbody>
    <body name="body1"><freejoint/><geom type="sphere" size=".1"/></body>
    <body name="body2"><freejoint/><geom type="sphere" size=".1"/></body>
    <body name="body3"><freejoint/><geom type="sphere" size=".1"/></body>
  </worldbody>
  <contact><exclude body1="body1" body2="body2"/></contact>
</mujoco>"""
  t = mio.derive_tables(mjcf.load_string(xml))
  pairid = np.asarray(t["nxn_pairid"]).reshape(-1, 2)
  assert pairid.shape[0] == 3
  np.testing.assert_array_equal(pairid[:, 0], [-2, -1, -1])


def test_min_friction(built):  # :949-975 zero friction is clamped to MJ_MINMU
  xml = """
<mujoco>
  <worldbody>
    <body><geom type="sphere" size=".1" friction="0 0 0"/><joint type="slide"/></body>
    <body><geom type="sphere" size=".1" friction="0 0 0"/><joint type="slide"/></body>
  </worldbody>
</mujoco>"""
  d = collide(mjcf.load_string(xml), [0.0, 0.1])
  assert int(d["ncon"][0]) == 1
  np.testing.assert_allclose(d["con_friction"][0, 0], 1e-5)


TET = np.array([[-1, 0, 0.1], [1, 0, 0.1], [0, 1, 0.1], [0, 0.5, 1.1]], dtype=np.float64)  # :688-712 tetrahedron 0.1 above the plane z = 0


def tet_desc(real):
  from tests.test_device_ccd_mesh_on_host import Desc, V

  keep = [np.ascontiguousarray(TET.astype(real)), np.zeros(3, real), np.eye(3, dtype=real).reshape(-1).copy(), np.zeros(3, real)]
  d = Desc()
  d.type, d.vertnum, d.polynum = 7, 4, 0
  d.vert, d.pos, d.mat, d.size = (a.ctypes.data_as(V) for a in keep)
  d.graph = None
  return d, keep


@pytest.mark.parametrize("real", [np.float64, np.float32])
def test_plane_tetrahedron_oracle(built, real):
  from oracle import orc
  from tests.test_device_ccd_mesh_on_host import V

  d, keep = tet_desc(real)
  lib = orc._lib(np.dtype(real).itemsize)
  lib.orc_plane_convex_desc.restype = None
  lib.orc_plane_convex_desc.argtypes = [V, V, V, V, V]
  dist = np.zeros(4, real); pos = np.zeros((4, 3), real)
  n, p = np.array([0, 0, 1], real), np.zeros(3, real)
  P = lambda a: a.ctypes.data_as(V)
  lib.orc_plane_convex_desc(P(n), P(p), ctypes.byref(d), P(dist), P(pos))
  assert (dist > 0.05).all()


def test_plane_tetrahedron_device_routine(hlib):
  from tests import test_device_ccd_mesh_on_host as H

  d, keep = tet_desc(np.float32)
  hlib.hplane_mesh.restype = None
  hlib.hplane_mesh.argtypes = [H.V] * 5
  dist = np.zeros(4, np.float32); pos = np.zeros((4, 3), np.float32)
  n, p = np.array([0, 0, 1], np.float32), np.zeros(3, np.float32)
  P = lambda a: a.ctypes.data_as(H.V)
  hlib.hplane_mesh(P(n), P(p), ctypes.byref(d), P(dist), P(pos))
  assert (dist > 0.05).all()
