"""Golden vectors produced by EXECUTING the reference's own device functions (unmodified sources from /root/reference)
under tools/warp_shim.py.  Run in the build container (where /root/reference exists):

  python tools/make_reference_goldens.py            # writes tests/golden/reference_colliders.json

The vectors pin the oracle's narrowphase (tests/test_oracle_golden_colliders.py) against the reference implementation
itself: /root/reference/mujoco_warp/_src/collision_primitive_core.py (all primitive pair functions) with math.py's
quat_to_mat / make_frame.  Scalars are evaluated in double precision, so the fp64 oracle must agree to ~1e-9.
"""

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import warp_shim  # noqa: E402

wp = warp_shim.install()
rmath = warp_shim.load_reference_module("math")
core = warp_shim.load_reference_module("collision_primitive_core")

# pair -> (type1, size1, type2, size2); sizes are MuJoCo geom_size triples
PAIRS = {
  "plane_sphere": ("plane", [0, 0, 0.05], "sphere", [0.11, 0, 0]),
  "plane_capsule": ("plane", [0, 0, 0.05], "capsule", [0.06, 0.17, 0]),
  "plane_ellipsoid": ("plane", [0, 0, 0.05], "ellipsoid", [0.1, 0.2, 0.05]),
  "plane_cylinder": ("plane", [0, 0, 0.05], "cylinder", [0.1, 0.2, 0]),
  "plane_box": ("plane", [0, 0, 0.05], "box", [0.2, 0.15, 0.1]),
  "sphere_sphere": ("sphere", [0.1, 0, 0], "sphere", [0.13, 0, 0]),
  "sphere_capsule": ("sphere", [0.1, 0, 0], "capsule", [0.06, 0.17, 0]),
  "sphere_cylinder": ("sphere", [0.07, 0, 0], "cylinder", [0.1, 0.2, 0]),
  "sphere_box": ("sphere", [0.1, 0, 0], "box", [0.2, 0.15, 0.1]),
  "capsule_capsule": ("capsule", [0.05, 0.2, 0], "capsule", [0.06, 0.17, 0]),
  "capsule_box": ("capsule", [0.05, 0.2, 0], "box", [0.2, 0.15, 0.1]),
  "box_box": ("box", [0.2, 0.15, 0.1], "box", [0.12, 0.1, 0.25]),
}
PLANE_POS = [0.1, -0.2, 0.05]
PLANE_QUAT = [0.9396926, 0.2, -0.25, 0.1]  # normalised below
MARGIN = 0.02  # sum of the two geoms' margins
NCASE = 48


def v3(a):
  return wp.vec3(float(a[0]), float(a[1]), float(a[2]))


def rand_quat(rng):
  q = rng.standard_normal(4)
  return q / np.linalg.norm(q)


def tolist(x):
  if isinstance(x, warp_shim.Vec):
    return [float(t) for t in x.v]
  if isinstance(x, warp_shim.Mat):
    return [[float(t) for t in r] for r in x.m]
  return float(x)


def run_pair(name, pos1, rot1, size1, pos2, rot2, size2):
  """-> (dist list, pos list, normal list) exactly as the reference wrapper consumes them (collision_primitive.py:281-1333)"""
  ax1 = wp.vec3(rot1[0, 2], rot1[1, 2], rot1[2, 2])
  ax2 = wp.vec3(rot2[0, 2], rot2[1, 2], rot2[2, 2])
  if name == "plane_sphere":
    d, p = core.plane_sphere(ax1, pos1, pos2, size2[0])
    return [d], [p], [ax1]
  if name == "plane_capsule":
    d, p, frame = core.plane_capsule(ax1, pos1, pos2, ax2, size2[0], size2[1])
    return [d[0], d[1]], [p[0], p[1]], frame  # shared frame (not make_frame)
  if name == "plane_ellipsoid":
    d, p, n = core.plane_ellipsoid(ax1, pos1, pos2, rot2, size2)
    return [d], [p], [n]
  if name == "plane_cylinder":
    d, p, n = core.plane_cylinder(ax1, pos1, pos2, ax2, size2[0], size2[1])
    return [d[i] for i in range(4)], [p[i] for i in range(4)], [n] * 4
  if name == "plane_box":
    d, p, n = core.plane_box(ax1, pos1, pos2, rot2, size2)
    return [d[i] for i in range(8)], [p[i] for i in range(8)], [n] * 8
  if name == "sphere_sphere":
    d, p, n = core.sphere_sphere(pos1, size1[0], pos2, size2[0])
    return [d], [p], [n]
  if name == "sphere_capsule":
    d, p, n = core.sphere_capsule(pos1, size1[0], pos2, ax2, size2[0], size2[1])
    return [d], [p], [n]
  if name == "sphere_cylinder":
    d, p, n = core.sphere_cylinder(pos1, size1[0], pos2, ax2, size2[0], size2[1])
    return [d], [p], [n]
  if name == "sphere_box":
    d, p, n = core.sphere_box(pos1, size1[0], pos2, rot2, size2)
    return [d], [p], [n]
  if name == "capsule_capsule":
    d, p, n = core.capsule_capsule(pos1, ax1, size1[0], size1[1], pos2, ax2, size2[0], size2[1], MARGIN)
    return [d[0], d[1]], [p[0], p[1]], [n[0], n[1]]
  if name == "capsule_box":
    d, p, n = core.capsule_box(pos1, ax1, size1[0], size1[1], pos2, rot2, size2)
    return [d[0], d[1]], [p[0], p[1]], [n[0], n[1]]
  if name == "box_box":
    d, p, n = core.box_box(pos1, rot1, size1, pos2, rot2, size2, MARGIN)
    return [d[i] for i in range(8)], [p[i] for i in range(8)], [n[i] for i in range(8)]
  raise KeyError(name)


def main():
  rng = np.random.default_rng(20260922)
  pq = np.array(PLANE_QUAT) / np.linalg.norm(PLANE_QUAT)
  out = {"margin": MARGIN, "plane_pos": PLANE_POS, "plane_quat": pq.tolist(), "pairs": {}}
  for name, (t1, s1, t2, s2) in PAIRS.items():
    cases = []
    r1 = 0.0 if t1 == "plane" else float(np.linalg.norm(s1) if t1 in ("box", "ellipsoid") else s1[0] + s1[1])
    r2 = float(np.linalg.norm(s2) if t2 in ("box", "ellipsoid") else s2[0] + s2[1])
    for c in range(NCASE):
      if t1 == "plane":
        p1, q1 = np.array(PLANE_POS), pq
        n = np.array(tolist(rmath.quat_to_mat(wp.quat(*q1))))[:, 2]
        lateral = rng.uniform(-0.5, 0.5, 3)
        p2 = p1 + lateral - n * (lateral @ n) + n * rng.uniform(-0.05, 1.5) * r2
      else:
        p1, q1 = rng.uniform(-0.1, 0.1, 3), rand_quat(rng)
        dirv = rng.standard_normal(3)
        p2 = p1 + dirv / np.linalg.norm(dirv) * rng.uniform(0.25, 1.25) * (r1 + r2)
      q2 = rand_quat(rng)
      if c % 8 == 7 and t1 != "plane":  # aligned configurations: parallel axes / face-face stacks
        q1 = np.array([1.0, 0, 0, 0]); q2 = np.array([1.0, 0, 0, 0])
        p2 = p1 + np.array([0.03 * (c // 8), 0.02, (s1[2] + s2[2] if t1 == "box" else 0.9 * (r1 + r2) * 0.5)])
      rot1, rot2 = rmath.quat_to_mat(wp.quat(*q1)), rmath.quat_to_mat(wp.quat(*q2))
      d, p, nrm = run_pair(name, v3(p1), rot1, v3(s1), v3(p2), rot2, v3(s2))
      if isinstance(nrm, warp_shim.Mat):
        frames = [tolist(nrm)] * len(d)
      else:
        frames = [tolist(rmath.make_frame(x)) for x in nrm]
      cases.append({"pos1": p1.tolist(), "quat1": np.asarray(q1).tolist(), "pos2": p2.tolist(), "quat2": q2.tolist(),
                    "dist": [tolist(x) for x in d], "pos": [tolist(x) for x in p], "frame": frames})
    out["pairs"][name] = {"type1": t1, "size1": s1, "type2": t2, "size2": s2, "cases": cases}
    nd = sum(1 for cs in cases for x in cs["dist"] if x < MARGIN)
    print(f"{name}: {len(cases)} cases, {nd} contacts within margin")
  path = os.path.join(ROOT, "tests", "golden", "reference_colliders.json")
  with open(path, "w") as f:
    json.dump(out, f)
  print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
  main()
