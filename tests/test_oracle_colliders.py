"""Geometry pins for the box / cylinder / ellipsoid primitive colliders of the oracle.

The reference validates these functions against live MuJoCo (collision_driver_test.py), which is not available offline, so each
collider is pinned here against an independent brute-force statement of the geometry it computes: point-to-box and
point-to-cylinder distances, dense sampling of a capsule's segment, a 15-axis separating-axis computation for box pairs,
corner heights for plane-box, rim sampling for plane-cylinder and the support function for plane-ellipsoid."""

import numpy as np
import pytest

from tests import util

XML = """
<mujoco>
  <option timestep="0.002"><flag nativeccd="disable"/></option>
  <worldbody>
    {plane}
    <body name="a" pos="0 0 1"><freejoint/><geom name="ga" type="{ta}" size="{sa}" margin="{margin}"/></body>
    {bodyb}
  </worldbody>
</mujoco>
"""
BODYB = '<body name="b" pos="1 0 1"><freejoint/><geom name="gb" type="{tb}" size="{sb}" margin="{margin}"/></body>'
PLANE = '<geom name="floor" type="plane" size="0 0 .05" margin="{margin}"/>'


def _scene(ta, sa, tb=None, sb=None, plane=False, margin=0.0):
  from mujoco_warp_b200._src import mjcf

  xml = XML.format(plane=PLANE.format(margin=margin) if plane else "", ta=ta, sa=sa, margin=margin,
                   bodyb=BODYB.format(tb=tb, sb=sb, margin=margin) if tb else "")
  return mjcf.load_string(xml)


def _rand_quat(rng):
  q = rng.standard_normal(4)
  return q / np.linalg.norm(q)


def _qmat(q):
  w, x, y, z = q
  return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                   [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                   [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _run(mjm, qpos, nconmax=16):
  o = util.make_oracle(mjm, len(qpos), nconmax, 8 * nconmax)
  o.set_state(qpos=np.asarray(qpos))
  o.forward()
  assert not (o.d["overflow"] & (1 << 30)).any(), "collider pair not restated"
  return o.d


def _point_box(p, bpos, R, size):
  """distance from a point to a solid box (0 inside) and the closest point"""
  loc = R.T @ (p - bpos)
  cl = np.clip(loc, -size, size)
  return np.linalg.norm(loc - cl), bpos + R @ cl


def test_sphere_box_distance(built):
  mjm = _scene("sphere", "0.1", "box", "0.2 0.15 0.1", margin=10.0)
  rng = np.random.default_rng(0)
  n = 64
  qpos = np.zeros((n, 14))
  for w in range(n):
    qpos[w, :3] = rng.uniform(-0.5, 0.5, 3); qpos[w, 3:7] = [1, 0, 0, 0]
    qpos[w, 7:10] = rng.uniform(-0.5, 0.5, 3); qpos[w, 10:14] = _rand_quat(rng)
  d = _run(mjm, qpos)
  size = np.array([0.2, 0.15, 0.1])
  for w in range(n):
    assert d["ncon"][w] == 1
    R = _qmat(qpos[w, 10:14])
    dist, cl = _point_box(qpos[w, :3], qpos[w, 7:10], R, size)
    if dist < 1e-9:
      continue  # centre inside the box: covered by the penetration test below
    np.testing.assert_allclose(d["con_dist"][w, 0], dist - 0.1, atol=1e-9)
    nrm = (cl - qpos[w, :3]) / dist
    np.testing.assert_allclose(d["con_frame"][w, 0, 0], nrm, atol=1e-9)
    np.testing.assert_allclose(d["con_pos"][w, 0], 0.5 * (cl + qpos[w, :3] + 0.1 * nrm), atol=1e-9)


def test_sphere_inside_box(built):
  mjm = _scene("sphere", "0.05", "box", "0.2 0.15 0.1")
  qpos = np.zeros((3, 14)); qpos[:, 3] = 1; qpos[:, 10] = 1
  qpos[0, :3] = [0.18, 0, 0]     # nearest face +x: pushed out along +x, so the sphere->box normal is -x
  qpos[1, :3] = [0, -0.14, 0.0]  # nearest face -y
  qpos[2, :3] = [0, 0, 0.09]     # nearest face +z
  d = _run(mjm, qpos)
  np.testing.assert_allclose(d["con_dist"][:, 0], [-0.02 - 0.05, -0.01 - 0.05, -0.01 - 0.05], atol=1e-12)
  np.testing.assert_allclose(d["con_frame"][:, 0, 0], [[-1, 0, 0], [0, 1, 0], [0, 0, -1]], atol=1e-12)


def test_capsule_box_first_contact_is_closest_point(built):
  rad, hl, size = 0.05, 0.2, np.array([0.2, 0.15, 0.1])
  mjm = _scene("capsule", f"{rad} {hl}", "box", "0.2 0.15 0.1", margin=10.0)
  rng = np.random.default_rng(1)
  n = 200
  qpos = np.zeros((n, 14))
  for w in range(n):
    qpos[w, :3] = rng.uniform(-0.5, 0.5, 3); qpos[w, 3:7] = _rand_quat(rng)
    qpos[w, 7:10] = rng.uniform(-0.2, 0.2, 3); qpos[w, 10:14] = _rand_quat(rng)
  d = _run(mjm, qpos)
  ts = np.linspace(-1, 1, 4001)
  checked = 0
  for w in range(n):
    Rc, Rb = _qmat(qpos[w, 3:7]), _qmat(qpos[w, 10:14])
    pts = qpos[w, :3] + np.outer(ts * hl, Rc[:, 2])
    loc = (pts - qpos[w, 7:10]) @ Rb
    dist = np.linalg.norm(loc - np.clip(loc, -size, size), axis=1)
    if dist.min() < 1e-3:
      continue  # segment touches the box: the closest feature is not unique
    checked += 1
    assert d["ncon"][w] >= 1
    ids = d["con_geomcollisionid"][w, : d["ncon"][w]]
    assert ids[0] == 0
    np.testing.assert_allclose(d["con_dist"][w, 0], dist.min() - rad, atol=2e-6)
    # the optional second contact is a sphere-box test at another segment point: never closer than the first
    if d["ncon"][w] == 2:
      assert d["con_dist"][w, 1] >= d["con_dist"][w, 0] - 1e-9
      p2 = d["con_pos"][w, 1] + d["con_frame"][w, 1, 0] * 0.5 * d["con_dist"][w, 1]  # point on the box surface
      assert _point_box(p2, qpos[w, 7:10], Rb, size)[0] < 1e-9
  assert checked > 100


def _sat(pos1, R1, s1, pos2, R2, s2):
  """largest signed separation over the 15 candidate axes (negative = overlap depth)"""
  axes = [R1[:, i] for i in range(3)] + [R2[:, i] for i in range(3)]
  for i in range(3):
    for j in range(3):
      c = np.cross(R1[:, i], R2[:, j])
      if np.linalg.norm(c) > 1e-6:
        axes.append(c / np.linalg.norm(c))
  best = -np.inf
  for a in axes:
    r1 = np.abs(R1.T @ a) @ s1
    r2 = np.abs(R2.T @ a) @ s2
    best = max(best, abs(a @ (pos2 - pos1)) - r1 - r2)
  return best


def test_box_box_depth_matches_separating_axis(built):
  s1, s2 = np.array([0.2, 0.15, 0.1]), np.array([0.12, 0.1, 0.25])
  margin = 0.02
  mjm = _scene("box", "0.2 0.15 0.1", "box", "0.12 0.1 0.25", margin=margin / 2)
  rng = np.random.default_rng(2)
  n = 3000
  qpos = np.zeros((n, 14))
  for w in range(n):
    qpos[w, :3] = rng.uniform(-0.1, 0.1, 3); qpos[w, 3:7] = _rand_quat(rng)
    qpos[w, 7:10] = qpos[w, :3] + rng.uniform(-0.45, 0.45, 3); qpos[w, 10:14] = _rand_quat(rng)
  # a few axis-aligned stacks (face-face, 4+ contacts)
  for w in range(8):
    qpos[w, :7] = [0, 0, 0, 1, 0, 0, 0]
    qpos[w, 7:14] = [0.03 * w, 0.01 * w, 0.1 + 0.25 - 0.004 * w, 1, 0, 0, 0]
  d = _run(mjm, qpos, nconmax=8)
  hit = exact = 0
  for w in range(n):
    R1, R2 = _qmat(qpos[w, 3:7]), _qmat(qpos[w, 10:14])
    sep = _sat(qpos[w, :3], R1, s1, qpos[w, 7:10], R2, s2)
    nc = d["ncon"][w]
    if sep > margin + 1e-9:
      assert nc == 0, (w, sep)
      continue
    if sep > -1e-4 or nc == 0 or sep < -0.03:
      # separated-within-margin pairs: the reference's edge-edge branch reports depths that differ from the axis
      # separation (checked bit-for-bit against the reference itself in test_oracle_golden_colliders.py);
      # deep overlaps: the clipped patch may reach below the least-penetration depth
      continue
    hit += 1
    dist = d["con_dist"][w, :nc]
    assert dist.min() <= sep + 1e-6, f"world {w}"  # never shallower than the least-penetration depth
    exact += abs(dist.min() - sep) < 1e-6  # ... and equal to it except in rare edge-edge configurations
    nrm = d["con_frame"][w, 0, 0]
    np.testing.assert_allclose(np.linalg.norm(nrm), 1, atol=1e-9)
    assert nrm @ (qpos[w, 7:10] - qpos[w, :3]) > 0  # from box 1 to box 2
    for k in range(nc):
      p = d["con_pos"][w, k]
      # the contact point sits midway between the two surfaces: within |dist|/2 (+ slack) of both solids
      assert _point_box(p, qpos[w, :3], R1, s1)[0] <= abs(dist[k]) / 2 + 1e-6
      assert _point_box(p, qpos[w, 7:10], R2, s2)[0] <= abs(dist[k]) / 2 + 1e-6
  assert hit > 100 and exact >= 0.97 * hit, (hit, exact)
  assert (d["ncon"][:8] >= 3).all(), "stacked boxes need a contact patch"


def test_plane_box_corners(built):
  mjm = _scene("box", "0.2 0.15 0.1", plane=True, margin=0.01)
  rng = np.random.default_rng(3)
  n = 64
  qpos = np.zeros((n, 7))
  for w in range(n):
    qpos[w, :3] = [rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(0.05, 0.3)]
    qpos[w, 3:7] = _rand_quat(rng) if w else [1, 0, 0, 0]
  d = _run(mjm, qpos)
  size = np.array([0.2, 0.15, 0.1])
  for w in range(n):
    R = _qmat(qpos[w, 3:7])
    corners = np.array([[(1 if i & 1 else -1), (1 if i & 2 else -1), (1 if i & 4 else -1)] for i in range(8)]) * size
    z = (corners @ R.T + qpos[w, :3])[:, 2]
    want = np.nonzero(z < 0.02)[0]  # margin of plane + box
    nc = d["ncon"][w]
    np.testing.assert_array_equal(d["con_geomcollisionid"][w, :nc], want)
    np.testing.assert_allclose(d["con_dist"][w, :nc], z[want], atol=1e-12)
    np.testing.assert_allclose(d["con_frame"][w, :nc, 0], np.tile([0, 0, 1.0], (nc, 1)), atol=1e-12)


def test_plane_cylinder_and_ellipsoid(built):
  rng = np.random.default_rng(4)
  n = 64
  qpos = np.zeros((n, 7))
  for w in range(n):
    qpos[w, :3] = [0, 0, rng.uniform(0.1, 0.4)]
    qpos[w, 3:7] = _rand_quat(rng)
  # cylinder: first contact = lowest rim point
  mjm = _scene("cylinder", "0.1 0.2", plane=True, margin=5.0)
  d = _run(mjm, qpos)
  th = np.linspace(0, 2 * np.pi, 20001)
  for w in range(n):
    R = _qmat(qpos[w, 3:7])
    rim = np.stack([0.1 * np.cos(th), 0.1 * np.sin(th), np.zeros_like(th)], 1)
    zmin = min(((rim + [0, 0, h]) @ R.T + qpos[w, :3])[:, 2].min() for h in (-0.2, 0.2))
    assert d["ncon"][w] == 4
    np.testing.assert_allclose(d["con_dist"][w, 0], zmin, atol=1e-7)
    # all four points lie on the cylinder surface (after undoing the half-distance shift along the normal)
    for k in range(4):
      p = d["con_pos"][w, k] + np.array([0, 0, 0.5 * d["con_dist"][w, k]])
      loc = R.T @ (p - qpos[w, :3])
      np.testing.assert_allclose(np.hypot(loc[0], loc[1]), 0.1, atol=1e-9)
      np.testing.assert_allclose(abs(loc[2]), 0.2, atol=1e-9)
  # ellipsoid: support function
  mjm = _scene("ellipsoid", "0.1 0.2 0.05", plane=True, margin=5.0)
  d = _run(mjm, qpos)
  size = np.array([0.1, 0.2, 0.05])
  for w in range(n):
    R = _qmat(qpos[w, 3:7])
    want = qpos[w, 2] - np.linalg.norm(size * (R.T @ np.array([0, 0, 1.0])))
    np.testing.assert_allclose(d["con_dist"][w, 0], want, atol=1e-12)


def test_sphere_cylinder_distance(built):
  mjm = _scene("sphere", "0.07", "cylinder", "0.1 0.2", margin=10.0)
  rng = np.random.default_rng(5)
  n = 128
  qpos = np.zeros((n, 14))
  for w in range(n):
    qpos[w, :3] = rng.uniform(-0.5, 0.5, 3); qpos[w, 3] = 1
    qpos[w, 7:10] = rng.uniform(-0.1, 0.1, 3); qpos[w, 10:14] = _rand_quat(rng)
  d = _run(mjm, qpos)
  checked = 0
  for w in range(n):
    R = _qmat(qpos[w, 10:14])
    loc = R.T @ (qpos[w, :3] - qpos[w, 7:10])
    rho, z = np.hypot(loc[0], loc[1]), abs(loc[2])
    dr, dz = rho - 0.1, z - 0.2
    if dr <= 0 and dz <= 0:
      continue  # centre inside the cylinder
    checked += 1
    want = np.hypot(max(dr, 0), max(dz, 0))
    np.testing.assert_allclose(d["con_dist"][w, 0], want - 0.07, atol=1e-9)
  assert checked > 100


def test_boxes_rest_on_plane_and_each_other(built):
  """Dynamics sanity through the whole step: a box dropped flat on the floor and a second box stacked on it (box-box
  primitive, nativeccd disabled) come to rest with sub-millimetre penetration."""
  from mujoco_warp_b200._src import mjcf

  mjm = mjcf.load_string("""
<mujoco>
  <option timestep="0.002"><flag nativeccd="disable"/></option>
  <worldbody>
    <geom type="plane" size="0 0 .05"/>
    <body pos="0 0 0.1"><freejoint/><geom type="box" size="0.2 0.2 0.1" density="500"/></body>
    <body pos="0.05 0.02 0.28"><freejoint/><geom type="box" size="0.1 0.1 0.08" density="500"/></body>
    <body pos="0.6 0 0.05" euler="90 0 0"><freejoint/><geom type="capsule" size="0.05 0.15"/></body>
    <body pos="0.6 0 0.18"><freejoint/><geom type="box" size="0.08 0.08 0.08" density="300"/></body>
  </worldbody>
</mujoco>""")
  o = util.make_oracle(mjm, 1, 32, 160)
  for _ in range(500):
    o.step()
  d = o.d
  assert d["overflow"][0] == 0
  np.testing.assert_allclose(d["qpos"][0, 2], 0.1, atol=2e-3)
  np.testing.assert_allclose(d["qpos"][0, 9], 0.28, atol=4e-3)
  assert np.abs(d["qvel"][0, :12]).max() < 5e-3
  assert np.isfinite(d["qpos"]).all()
