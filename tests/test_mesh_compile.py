"""Host-side mesh preparation (mujoco_warp_b200/_src/mesh.py): hull, graph, polygons and mass properties of known solids."""
import numpy as np
import pytest

from mujoco_warp_b200._src import mesh


def cube(h=(0.5, 0.3, 0.2), extra=True):
  c = np.array([[sx * h[0], sy * h[1], sz * h[2]] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=float)
  if extra:  # interior points and points on faces / edges must not change the hull polygons
    c = np.vstack([c, [[0, 0, 0], [0.1, 0.05, -0.1], [h[0], 0, 0], [0, h[1], h[2]]]])
  return c


def test_box_mesh_tables():
  pts = cube() + np.array([0.3, -0.2, 0.7])
  out = mesh.process(pts)
  np.testing.assert_allclose(out["volume"], 8 * 0.5 * 0.3 * 0.2, rtol=1e-12)
  np.testing.assert_allclose(out["pos"], [0.3, -0.2, 0.7], atol=1e-12)
  m = out["volume"]
  want = np.sort([m / 3 * (0.3**2 + 0.2**2), m / 3 * (0.5**2 + 0.2**2), m / 3 * (0.5**2 + 0.3**2)])[::-1]
  np.testing.assert_allclose(out["inertia"], want, rtol=1e-12)  # principal moments, descending
  # six quadrilateral polygons with outward unit normals along the principal axes
  assert len(out["polynormal"]) == 6 and (out["polyvertnum"] == 4).all()
  np.testing.assert_allclose(np.abs(out["polynormal"]).max(axis=1), 1.0, atol=1e-12)
  adr = np.concatenate(([0], np.cumsum(out["polyvertnum"])))
  for p in range(6):
    loop = out["vert"][out["polyvert"][adr[p] : adr[p + 1]]]
    n = np.cross(loop[1] - loop[0], loop[2] - loop[1])
    assert np.dot(n, out["polynormal"][p]) > 0  # counter-clockwise seen from outside
    assert np.allclose((loop - loop[0]) @ out["polynormal"][p], 0, atol=1e-12)  # planar
  # every corner belongs to three polygons, the extra points to none
  assert sorted(out["polymapnum"].tolist()) == [0] * 4 + [3] * 8
  # hull graph: 8 vertices, 12 triangles, symmetric neighbour lists terminated by -1
  g = out["graph"]
  nv, nf = int(g[0]), int(g[1])
  assert (nv, nf) == (8, 12) and len(g) == 2 + 3 * nv + 6 * nf
  edgeadr, glob, edges = g[2 : 2 + nv], g[2 + nv : 2 + 2 * nv], g[2 + 2 * nv : 2 + 3 * nv + 3 * nf]
  nbrs = []
  for i in range(nv):
    j, lst = edgeadr[i], []
    while edges[j] >= 0:
      lst.append(int(edges[j])); j += 1
    nbrs.append(lst)
  assert all(i in nbrs[j] for i in range(nv) for j in nbrs[i])
  assert sum(len(x) for x in nbrs) == 3 * nf  # 2 * (3 nf / 2) directed edges
  # hill climbing over the graph finds the support vertex of any direction (what collision_gjk.py:171-194 does)
  rng = np.random.default_rng(0)
  for _ in range(50):
    d = rng.normal(size=3)
    i, prev = 0, -1
    while i != prev:
      prev = i
      for j in nbrs[prev]:
        if out["vert"][glob[j]] @ d > out["vert"][glob[i]] @ d:
          i = j
    assert np.isclose(out["vert"][glob[i]] @ d, (out["vert"] @ d).max())


def test_tetrahedron_and_random_cloud():
  tet = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=float)
  out = mesh.process(tet)
  np.testing.assert_allclose(out["volume"], 1 / 6, rtol=1e-12)
  np.testing.assert_allclose(out["pos"], [0.25, 0.25, 0.25], atol=1e-12)
  assert len(out["polynormal"]) == 4 and (out["polyvertnum"] == 3).all()
  rng = np.random.default_rng(1)
  pts = rng.normal(size=(200, 3)) * [1.0, 0.6, 0.3]
  out = mesh.process(pts, scale=(2, 2, 2))
  g = out["graph"]
  nv, nf = int(g[0]), int(g[1])
  assert nv - 3 * nf // 2 + nf == 2  # Euler characteristic of a closed triangulated surface
  np.testing.assert_allclose(out["vert"][np.unique(g[2 + nv : 2 + 2 * nv])].mean(axis=0), out["vert"][g[2 + nv : 2 + 2 * nv]].mean(axis=0))
  assert out["rbound"] >= np.abs(out["aabb_size"]).max() - 1e-12
  # the mesh frame is centred on the centre of mass with principal axes: products of inertia vanish
  _, com, inertia = mesh.mass_properties(out["vert"], mesh.convex_hull(out["vert"])[1])
  np.testing.assert_allclose(com, 0, atol=1e-9)
  assert np.abs(inertia - np.diag(np.diag(inertia))).max() < 1e-9 * np.abs(inertia).max()


MESH_XML = """<mujoco><asset>
 <mesh name="cube" vertex="-1 -1 -1  1 -1 -1  -1 1 -1  1 1 -1  -1 -1 1  1 -1 1  -1 1 1  1 1 1" scale="0.1 0.05 0.02"/>
</asset><worldbody>
 <geom type="plane" size="0 0 .05"/>
 <body pos="1 0 1"><freejoint/><geom type="mesh" mesh="cube" pos="0.1 0 0" euler="0 0 30" {col}/></body>
 <body pos="2 0 1"><freejoint/><geom type="box" size="0.1 0.05 0.02" pos="0.1 0 0" euler="0 0 30" {col}/></body>
</worldbody></mujoco>"""


def test_mesh_geom_compiles_like_the_equivalent_box():
  """A cube given as a mesh asset yields the mass, inertia, centre of mass and bounds of the same box primitive; the mesh tables
  of the Model are filled; colliding mesh geoms are routed to the convex pass of the mesh build of the collision kernel."""
  from mujoco_warp_b200._src import io as mio
  from mujoco_warp_b200._src import mjcf

  m = mjcf.load_string(MESH_XML.format(col='contype="0" conaffinity="0"'))
  assert m.nmesh == 1 and list(m.geom_dataid) == [-1, 0, -1] and m.geom_type[1] == 7
  np.testing.assert_allclose(m.body_mass[1], m.body_mass[2], rtol=1e-12)
  np.testing.assert_allclose(m.body_inertia[1], m.body_inertia[2], rtol=1e-10)
  np.testing.assert_allclose(m.body_ipos[1], m.body_ipos[2], atol=1e-12)
  np.testing.assert_allclose(np.sort(m.geom_size[1]), np.sort(m.geom_size[2]), rtol=1e-12)
  np.testing.assert_allclose(m.geom_rbound[1], m.geom_rbound[2], rtol=1e-12)
  assert m.mesh_vertnum[0] == 8 and m.mesh_polynum[0] == 6 and m.mesh_graph[0] == 8 and len(m.mesh_polymap) == 24
  # world-frame corners of the mesh geom coincide with the box's corners
  kin = mjcf.kinematics_np(m, m.qpos0)
  gx, gR = np.asarray(kin.geom_xpos).reshape(-1, 3), np.asarray(kin.geom_xmat).reshape(-1, 3, 3)
  mesh_corners = (gR[1] @ m.mesh_vert.T).T + gx[1] - [1, 0, 0]
  s = m.geom_size[2]
  box_corners = np.array([[sx * s[0], sy * s[1], sz * s[2]] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)])
  box_corners = (gR[2] @ box_corners.T).T + gx[2] - [2, 0, 0]
  d = np.linalg.norm(mesh_corners[:, None] - box_corners[None], axis=2)
  assert d.min(axis=1).max() < 1e-12 and d.min(axis=0).max() < 1e-12
  assert mio.derive_tables(m)["has_convex_pair"] == 0  # visual only
  assert mio.derive_tables(mjcf.load_string(MESH_XML.format(col="")))["has_convex_pair"] == 1  # box - mesh goes through GJK / EPA
