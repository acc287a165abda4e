"""The reference's known-answer tests for mesh geoms in the convex path (collision_gjk_test.py:341 box_mesh_distance, :405 mesh_mesh_contact,
:441 mesh_mesh_contact2, :528 mesh_mesh_ccd, :648 sphere_mesh_margin), transcribed as data.  Meshes are compiled by this repo's MJCF
compiler (hull, hull graph, polygons), posed by its kinematics and run through the oracle's convex pair routine and through the device
routine (CCD_MESH build of csrc/mjb_ccd.cuh) compiled for the host."""
import ctypes

import numpy as np
import pytest

from mujoco_warp_b200._src import mjcf
from oracle import orc
from tests.test_device_ccd_mesh_on_host import I, V, hlib, make_desc  # noqa: F401  (hlib: fixture)

CUBE = "-1 -1 -1 1 -1 -1 1 1 -1 1 1 1 1 -1 1 -1 1 -1 -1 1 1 -1 -1 1"
CASES = {
  "box_mesh_distance": dict(  # :341
    asset=f'<mesh name="smallbox" scale="0.1 0.1 0.1" vertex="{CUBE}"/>',
    body='<geom pos="0 0 .90" type="box" size="0.5 0.5 0.1"/><geom pos="0 0 1.2" type="mesh" mesh="smallbox"/>', dist=0.1),
  "mesh_mesh_contact": dict(  # :405
    asset=f'<mesh name="box" scale=".5 .5 .1" vertex="{CUBE}"/><mesh name="smallbox" scale=".1 .1 .1" vertex="{CUBE}"/>',
    body='<geom pos="0 0 .09" type="mesh" mesh="smallbox"/><geom pos="0 0 -.1" type="mesh" mesh="box"/>', dist=-0.01),
  "mesh_mesh_contact2": dict(  # :441 degenerate geometry
    asset='<mesh name="mesh" vertex="-0.0611590669 -0.13801524 -0.158372656  0.0620514415 0.135089189 -0.159879193  -0.105518319 -0.100999095 -0.188289702 '
          '-0.107238553 -0.102976903 0.1569262  -0.0851279497 -0.122304708 0.156887323  -0.0590926372 -0.104567274 -0.242715642"/>',
    body='<geom name="geom1" type="mesh" mesh="mesh" pos="-0.141666584 0 0" quat="0.5425650813 0.0029009761 0.0001424328 0.8400087479"/>'
         '<geom name="geom2" type="mesh" mesh="mesh" pos="0.141666584 0 0" quat="0.5425650813 0.0029009761 0.0001424328 0.8400087479"/>',
    dist=-0.0031312597856874586, ncon=1),
  "mesh_mesh_ccd": dict(  # :528
    asset=f'<mesh name="smallbox" vertex="{CUBE}"/>',
    body='<geom pos="0 0 2" type="mesh" name="box1" mesh="smallbox"/><geom pos="0 1 3.99" euler="0 0 40" type="mesh" name="box2" mesh="smallbox"/>',
    multiccd=True, ncon=4),
  "sphere_mesh_margin": dict(  # :648
    asset=f'<mesh name="box" scale=".2 .2 .2" vertex="{CUBE}"/>',
    body='<geom type="sphere" pos="0 0 .349" size=".1"/><geom type="mesh" mesh="box"/>', margin=0.05, dist=-0.001),
}


def run(case, which, real, hlib=None):
  mjm = mjcf.load_string(f"<mujoco><asset>{case['asset']}</asset><worldbody>{case['body']}</worldbody></mujoco>")
  kin = mjcf.kinematics_np(mjm, mjm.qpos0)
  xpos, xmat = np.asarray(kin.geom_xpos, dtype=np.float64).reshape(-1, 3), np.asarray(kin.geom_xmat, dtype=np.float64).reshape(-1, 9)
  d1, k1 = make_desc(mjm, 0, xpos[0], xmat[0], real)
  d2, k2 = make_desc(mjm, 1, xpos[1], xmat[1], real)
  dist = np.zeros(1, real); w1 = np.zeros((4, 3), real); w2 = np.zeros((4, 3), real); ovf = np.zeros(1, np.int32)
  P = lambda a: a.ctypes.data_as(V)
  margin, multi = case.get("margin", 0.0), int(case.get("multiccd", False))
  if which == "device":
    hlib.hccd_desc.restype = ctypes.c_int
    n = hlib.hccd_desc(ctypes.byref(d1), ctypes.byref(d2), margin, 1e-6, 1e30, 35, 35, P(dist), P(w1), P(w2), P(ovf))
  else:
    lib = orc._lib(np.dtype(real).itemsize)
    c_real = ctypes.c_double if real is np.float64 else ctypes.c_float
    lib.orc_ccd_desc.restype = ctypes.c_int
    lib.orc_ccd_desc.argtypes = [V, V, c_real, c_real, c_real, I, I, I, V, V, V, V]
    n = lib.orc_ccd_desc(ctypes.byref(d1), ctypes.byref(d2), margin, 1e-6, 1e30, 35, 35, multi, P(dist), P(w1), P(w2), P(ovf))
  assert ovf[0] == 0
  return float(dist[0]), int(n)


def check(case, dist, ncon, multi_always=False):
  if "dist" in case:
    assert abs(dist - case["dist"]) < 0.5e-7, (dist, case["dist"])  # assertAlmostEqual, 7 places
  if "ncon" in case and not (multi_always and not case.get("multiccd", False) and case["ncon"] == 1 and ncon >= 1):
    assert ncon == case["ncon"], (ncon, case["ncon"])


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("real", [np.float64, np.float32])
def test_oracle_meets_reference_mesh_vectors(name, real):
  dist, ncon = run(CASES[name], "oracle", real)
  check(CASES[name], dist, ncon)


@pytest.mark.parametrize("name", sorted(CASES))
def test_device_routine_meets_reference_mesh_vectors(hlib, name):
  dist, ncon = run(CASES[name], "device", np.float32, hlib)
  check(CASES[name], dist, ncon, multi_always=True)  # the device routine always recovers the contact patch of box / mesh pairs
