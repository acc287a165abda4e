"""Pins on this repo's MJCF compiler (`_src/mjcf.py`), which stands in for MuJoCo's C compiler (absent here): "model compile
unpinned" is the stated status of the parity claim, so these tests hold the compiled benchmark models to everything that can be
checked without MuJoCo itself --

* the model facts the reference documents (benchmarks/humanoid/README.md:11-21,31-41, unitree_g1/README.md:13-23): body / dof /
  actuator / geom counts, timestep, solver, cone, integrator, matrix format;
* derived constants against an INDEPENDENT evaluation: the joint-space inertia at qpos0 assembled from finite-difference body
  Jacobians of the forward kinematics (no CRB recursion, none of the compiler's inertia code), then
  meaninertia = trace(M) / nv, dof_invweight0 = diag(M^-1) (averaged per ball / free-joint block), body_invweight0 =
  trace(J M^-1 J^T) / 3 for the translational and rotational Jacobians at the body's inertial frame (MuJoCo's set_const definitions);
* mass bookkeeping (subtree masses), left / right symmetry of the humanoid's limbs;
* geom bounding data against brute-force support sampling: no surface point lies outside geom_rbound / geom_aabb, and both are tight.
"""
import numpy as np
import pytest

from mujoco_warp_b200 import scenes
from mujoco_warp_b200._src import constants as C
from mujoco_warp_b200._src import io as mio
from mujoco_warp_b200._src import mjcf

FACTS = {  # README tables of the reference
  "humanoid": dict(path=scenes.HUMANOID, nbody=17, nv=27, nu=21, ngeom=20, timestep=0.005, solver=C.SOL_NEWTON, cone=C.CONE_PYRAMIDAL, integrator=C.INT_EULER, sparse=False),
  "three_humanoids": dict(path=scenes.THREE_HUMANOIDS, nbody=49, nv=81, nu=63, ngeom=58, timestep=0.005, solver=C.SOL_NEWTON, cone=C.CONE_PYRAMIDAL, integrator=C.INT_EULER, sparse=True),
  "unitree_g1_flat": dict(path=scenes.G1, nbody=31, nv=35, nu=29, ngeom=69, timestep=0.005, solver=C.SOL_NEWTON, cone=C.CONE_PYRAMIDAL, integrator=C.INT_IMPLICITFAST, sparse=True),
}


@pytest.fixture(scope="module", params=sorted(FACTS))
def model(request):
  return request.param, mjcf.load_any(FACTS[request.param]["path"])


def test_documented_model_facts(model):
  name, m = model
  f = FACTS[name]
  # the G1's 35 visual mesh geoms (STL files outside the tree, contype 0) are skipped by the compiler and counted
  assert m.nbody == f["nbody"] and m.nv == f["nv"] and m.nu == f["nu"] and m.ngeom + int(getattr(m, "skipped_mesh_geoms", 0)) == f["ngeom"]
  assert abs(m.opt.timestep - f["timestep"]) < 1e-12 and m.opt.solver == f["solver"] and m.opt.cone == f["cone"] and m.opt.integrator == f["integrator"]
  assert mio.is_sparse(m) == f["sparse"]


def _quat_mul(a, b):
  return np.array([a[0] * b[0] - a[1:] @ b[1:], *(a[0] * b[1:] + b[0] * a[1:] + np.cross(a[1:], b[1:]))])


def _perturb(m, qpos, dof, eps):
  """qpos moved by eps along one dof (MuJoCo's tangent-space convention: free / ball rotations are body-frame angular velocities)."""
  q = qpos.copy()
  j = int(m.dof_jntid[dof])
  k = dof - int(m.jnt_dofadr[j])
  qa, t = int(m.jnt_qposadr[j]), int(m.jnt_type[j])
  if t in (C.JNT_SLIDE, C.JNT_HINGE) or (t == C.JNT_FREE and k < 3):
    q[qa + k] += eps
    return q
  off, ax = (3, k - 3) if t == C.JNT_FREE else (0, k)
  w = np.zeros(3)
  w[ax] = eps
  dq = np.array([np.cos(eps / 2), *(np.sin(eps / 2) * w / eps)])
  q[qa + off : qa + off + 4] = _quat_mul(q[qa + off : qa + off + 4], dq)
  return q


def _fd_jacobians(m):
  """(nbody, 3, nv) translational (at the inertial frame origin) and rotational body Jacobians at qpos0, central differences of FK."""
  nb, nv, eps = m.nbody, m.nv, 1e-6
  q0 = np.asarray(m.qpos0, dtype=np.float64)
  Jp, Jr = np.zeros((nb, 3, nv)), np.zeros((nb, 3, nv))
  for d in range(nv):
    kp, km = mjcf.kinematics_np(m, _perturb(m, q0, d, eps)), mjcf.kinematics_np(m, _perturb(m, q0, d, -eps))
    Jp[:, :, d] = (np.asarray(kp.xipos) - np.asarray(km.xipos)) / (2 * eps)
    for b in range(nb):
      dR = np.asarray(kp.ximat[b]).reshape(3, 3) @ np.asarray(km.ximat[b]).reshape(3, 3).T  # R(+) R(-)^T = exp([w] 2 eps)
      Jr[b, :, d] = np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / (4 * eps)
  return Jp, Jr


def test_inertia_constants_against_an_independent_evaluation(model):
  name, m = model
  Jp, Jr = _fd_jacobians(m)
  kin = mjcf.kinematics_np(m, np.asarray(m.qpos0, dtype=np.float64))
  nv = m.nv
  M = np.diag(np.asarray(m.dof_armature, dtype=np.float64))
  for b in range(1, m.nbody):
    R = np.asarray(kin.ximat[b]).reshape(3, 3)
    M += m.body_mass[b] * Jp[b].T @ Jp[b] + Jr[b].T @ (R @ np.diag(m.body_inertia[b]) @ R.T) @ Jr[b]
  assert np.allclose(M, M.T, atol=1e-6) and np.linalg.eigvalsh(M).min() > 0
  np.testing.assert_allclose(m.stat.meaninertia, np.trace(M) / nv, rtol=1e-5)
  Minv = np.linalg.inv(M)
  dg = np.diag(Minv).copy()
  want = np.zeros(nv)
  for j in range(m.njnt):
    d, t = int(m.jnt_dofadr[j]), int(m.jnt_type[j])
    if t == C.JNT_FREE:
      want[d : d + 3], want[d + 3 : d + 6] = dg[d : d + 3].mean(), dg[d + 3 : d + 6].mean()
    elif t == C.JNT_BALL:
      want[d : d + 3] = dg[d : d + 3].mean()
    else:
      want[d] = dg[d]
  np.testing.assert_allclose(m.dof_invweight0, want, rtol=2e-4, atol=1e-9)
  for b in range(1, m.nbody):
    np.testing.assert_allclose(m.body_invweight0[b, 0], np.trace(Jp[b] @ Minv @ Jp[b].T) / 3, rtol=2e-4, atol=1e-9, err_msg=f"body {b} translational")
    np.testing.assert_allclose(m.body_invweight0[b, 1], np.trace(Jr[b] @ Minv @ Jr[b].T) / 3, rtol=2e-4, atol=1e-9, err_msg=f"body {b} rotational")
  # the CSR sparsity the compiler emits for M is exactly the set of (dof, ancestor dof) pairs, and covers M's nonzeros
  pattern = np.zeros((nv, nv), dtype=bool)
  for i in range(nv):
    k = i
    while k >= 0:
      pattern[i, k] = pattern[k, i] = True
      k = int(m.dof_parentid[k])
  assert (np.abs(M)[~pattern] < 1e-6 * np.abs(M).max()).all()
  rows = np.repeat(np.arange(nv), np.asarray(m.M_rownnz))
  assert len(rows) == len(m.M_colind) and pattern[rows, np.asarray(m.M_colind)].all() and len(rows) == int(np.tril(pattern).sum())


def test_mass_bookkeeping_and_symmetry(model):
  name, m = model
  sub = np.asarray(m.body_mass, dtype=np.float64).copy()
  for b in range(m.nbody - 1, 0, -1):
    sub[m.body_parentid[b]] += sub[b]
  np.testing.assert_allclose(m.body_subtreemass, sub, rtol=1e-12)
  assert (np.asarray(m.body_mass)[1:] > 0).all() and (np.asarray(m.body_inertia)[1:] > 0).all()
  # principal inertias satisfy the triangle inequality (a physical rigid body)
  I = np.sort(np.asarray(m.body_inertia)[1:], axis=1)
  assert (I[:, 0] + I[:, 1] >= I[:, 2] * (1 - 1e-9)).all()
  if name == "humanoid":
    names = list(m.names.body)
    for left in [n for n in names if n.endswith("_left")]:
      right = left[: -len("_left")] + "_right"
      if right in names:
        a, b = names.index(left), names.index(right)
        np.testing.assert_allclose(m.body_mass[a], m.body_mass[b], rtol=1e-12, err_msg=left)
        np.testing.assert_allclose(np.sort(m.body_inertia[a]), np.sort(m.body_inertia[b]), rtol=1e-9, err_msg=left)
        np.testing.assert_allclose(np.abs(m.body_ipos[a]), np.abs(m.body_ipos[b]), atol=1e-12, err_msg=left)


def _surface_samples(gtype, size, n=4000, seed=0):
  """Points on the surface of a primitive geom in its own frame (support points of random directions)."""
  rng = np.random.default_rng(seed)
  dirs = rng.normal(size=(n, 3))
  dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
  if gtype == C.GEOM_SPHERE:
    return dirs * size[0]
  if gtype == C.GEOM_CAPSULE:
    return dirs * size[0] + np.c_[np.zeros((n, 2)), np.sign(dirs[:, 2]) * size[1]]
  if gtype == C.GEOM_ELLIPSOID:
    s = dirs * size
    return s / np.linalg.norm(s, axis=1, keepdims=True) * size
  if gtype == C.GEOM_CYLINDER:
    r = np.linalg.norm(dirs[:, :2], axis=1, keepdims=True)
    return np.c_[dirs[:, :2] / np.maximum(r, 1e-12) * size[0], np.sign(dirs[:, 2]) * size[1]]
  if gtype == C.GEOM_BOX:
    return np.sign(dirs) * size
  return None


def test_geom_bounds_against_support_sampling(model):
  name, m = model
  checked = 0
  for g in range(m.ngeom):
    pts = _surface_samples(int(m.geom_type[g]), np.asarray(m.geom_size[g], dtype=np.float64))
    if pts is None:
      continue
    checked += 1
    r = np.linalg.norm(pts, axis=1).max()
    assert m.geom_rbound[g] >= r * (1 - 1e-9) and m.geom_rbound[g] <= r * 1.02 + 1e-9, (g, m.geom_rbound[g], r)
    c, h = np.asarray(m.geom_aabb[g]).reshape(2, 3)
    assert (np.abs(pts - c) <= h * (1 + 1e-9) + 1e-12).all(), g
    assert (np.abs(pts - c).max(axis=0) >= h * 0.98 - 1e-9).all(), g
  assert checked >= m.ngeom - 2


def test_mesh_files_obj_and_stl_compile_like_inline_vertices(tmp_path):
  """<mesh file=...> (OBJ, binary STL, ASCII STL under <compiler meshdir>) gives the same asset tables as the same vertices inline;
  a default-class scale applies; a missing file only breaks the model when a colliding geom needs it."""
  import struct

  from mujoco_warp_b200._src import mjcf

  v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [0.3, 0.3, 0.1], [1, 1, 1]], dtype=np.float64)
  from mujoco_warp_b200._src.mesh import convex_hull

  _, tris = convex_hull(v)
  d = tmp_path / "assets"
  d.mkdir()
  (d / "w.obj").write_text("".join(f"v {a} {b} {c}\n" for a, b, c in v) + "".join(f"f {a + 1}/1/1 {b + 1}/1/1 {c + 1}/1/1\n" for a, b, c in tris))
  with open(d / "w_bin.stl", "wb") as f:
    f.write(b"\0" * 80 + struct.pack("<I", len(tris)))
    for t in tris:
      f.write(struct.pack("<12fH", 0, 0, 0, *v[t].reshape(-1), 0))
  (d / "w_ascii.stl").write_text("solid w\n" + "".join("facet normal 0 0 0\nouter loop\n" + "".join(f"vertex {a} {b} {c}\n" for a, b, c in v[t]) + "endloop\nendfacet\n" for t in tris) + "endsolid w\n")
  body = '<worldbody><geom type="plane" size="0 0 .05"/><body pos="0 0 0.2"><freejoint/><geom type="mesh" mesh="{}"/></body></worldbody>'
  inline = mjcf.load_string(f'<mujoco><asset><mesh name="w" vertex="{" ".join(str(x) for x in v.reshape(-1))}" scale="0.1 0.1 0.1"/></asset>{body.format("w")}</mujoco>')
  for fn in ("w.obj", "w_bin.stl", "w_ascii.stl"):
    p = tmp_path / f"m_{fn}.xml"
    p.write_text(f'<mujoco><compiler meshdir="assets"/><default><default class="c"><mesh scale="0.1 0.1 0.1"/></default></default>'
                 f'<asset><mesh name="w" class="c" file="{fn}"/></asset>{body.format("w")}</mujoco>')
    m = mjcf.load(str(p))
    assert m.nmesh == 1 and int(m.mesh_vertnum[0]) == (7 if fn.endswith("obj") else len(np.unique(v[np.unique(tris)], axis=0)))
    np.testing.assert_allclose(m.body_mass, inline.body_mass, rtol=1e-9)
    np.testing.assert_allclose(m.body_inertia, inline.body_inertia, rtol=1e-8, atol=1e-14)
    np.testing.assert_allclose(m.geom_rbound, inline.geom_rbound, rtol=1e-9)
    np.testing.assert_allclose(np.sort(m.mesh_polynormal, axis=0), np.sort(inline.mesh_polynormal, axis=0), atol=1e-9)
  p = tmp_path / "missing.xml"
  p.write_text(f'<mujoco><compiler meshdir="assets"/><asset><mesh name="w" file="nope.obj"/></asset>{body.format("w")}</mujoco>')
  with pytest.raises(NotImplementedError):
    mjcf.load(str(p))
