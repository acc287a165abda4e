"""Mesh assets for the MJCF subset compiler: convex hull, hull graph, hull polygons and mass properties.

Produces the per-mesh tables the reference's Model carries (types.py:1214-1235): `mesh_vert` (asset vertices re-expressed in the
mesh frame: centre of mass at the origin, principal axes of inertia as coordinate axes, as MuJoCo's compiler does, with the
applied transform in `mesh_pos` / `mesh_quat`), `mesh_face`, `mesh_graph` (MuJoCo's convex-hull graph: numvert, numface,
vert_edgeadr[numvert], vert_globalid[numvert], edge_localid[numvert + 3 numface] with -1-terminated neighbour lists,
face_globalid[3 numface] -- the layout collision_gjk.py:171-194 hill-climbs over) and the hull polygons `mesh_poly*` /
`mesh_polymap*` that multi-contact recovery reads (collision_gjk.py:1585-1699, :1891-1912).

Host-side preparation only: the collision kernels do not take mesh geoms yet (put_model still refuses them); this is the
compiler half of that work, covered by tests/test_mesh_compile.py."""

from __future__ import annotations

import numpy as np


def convex_hull(points: np.ndarray):
  """(hull vertex ids ascending, outward-oriented triangles as global vertex ids) of a point cloud."""
  from scipy.spatial import ConvexHull

  pts = np.asarray(points, dtype=np.float64)
  hull = ConvexHull(pts, qhull_options="Qt")
  tris = hull.simplices.copy()
  for k, (tri, eq) in enumerate(zip(tris, hull.equations)):  # eq[:3] is the outward unit normal
    a, b, c = pts[tri]
    if np.dot(np.cross(b - a, c - a), eq[:3]) < 0:
      tris[k] = tri[[0, 2, 1]]
  return np.unique(tris), tris


def hull_graph(hull_ids: np.ndarray, tris: np.ndarray) -> np.ndarray:
  """MuJoCo's convex-hull graph as one int array (see module docstring)."""
  local = {int(g): i for i, g in enumerate(hull_ids)}
  nv, nf = len(hull_ids), len(tris)
  nbr = [[] for _ in range(nv)]
  for tri in tris:
    for a, b in ((0, 1), (1, 2), (2, 0)):
      i, j = local[int(tri[a])], local[int(tri[b])]
      if j not in nbr[i]:
        nbr[i].append(j)
      if i not in nbr[j]:
        nbr[j].append(i)
  edgeadr, edges = [], []
  for i in range(nv):
    edgeadr.append(len(edges))
    edges += sorted(nbr[i]) + [-1]
  assert len(edges) == nv + 3 * nf, "a closed triangulated hull has 3 nf / 2 edges, i.e. 3 nf directed neighbours plus nv terminators"
  face_local = [local[int(v)] for tri in tris for v in tri]
  return np.array([nv, nf] + edgeadr + [int(g) for g in hull_ids] + edges + face_local, dtype=np.int32)


def hull_polygons(points: np.ndarray, tris: np.ndarray, tol: float = 1e-6):
  """Merge coplanar neighbouring hull triangles into convex polygons: list of (unit outward normal, vertex loop counter-clockwise
  seen from outside, global vertex ids)."""
  pts = np.asarray(points, dtype=np.float64)
  normals = []
  for tri in tris:
    a, b, c = pts[tri]
    n = np.cross(b - a, c - a)
    normals.append(n / np.linalg.norm(n))
  normals = np.array(normals)
  edge_tri = {}
  for t, tri in enumerate(tris):
    for a, b in ((0, 1), (1, 2), (2, 0)):
      edge_tri.setdefault((int(tri[a]), int(tri[b])), t)
  # union-find over triangles that share an edge and a plane
  parent = list(range(len(tris)))

  def find(x):
    while parent[x] != x:
      parent[x] = parent[parent[x]]
      x = parent[x]
    return x

  for (a, b), t in edge_tri.items():
    u = edge_tri.get((b, a))
    if u is not None and np.dot(normals[t], normals[u]) > 1.0 - tol:
      parent[find(t)] = find(u)
  groups = {}
  for t in range(len(tris)):
    groups.setdefault(find(t), []).append(t)
  polys = []
  for members in groups.values():
    # boundary = directed edges whose reverse is not inside the group; chain them into one loop
    inside = {(int(tris[t][a]), int(tris[t][b])) for t in members for a, b in ((0, 1), (1, 2), (2, 0))}
    nxt = {a: b for (a, b) in inside if (b, a) not in inside}
    start = min(nxt)
    loop, v = [start], nxt[start]
    while v != start:
      loop.append(v)
      v = nxt[v]
    n = normals[members].mean(axis=0)
    n /= np.linalg.norm(n)
    # drop collinear vertices left over from the triangulation of a flat face
    keep = []
    for i, v in enumerate(loop):
      p, q, r = pts[loop[i - 1]], pts[v], pts[loop[(i + 1) % len(loop)]]
      if np.linalg.norm(np.cross(q - p, r - q)) > tol * np.linalg.norm(q - p) * np.linalg.norm(r - q):  # sine of the turn angle
        keep.append(v)
    polys.append((n, keep if len(keep) >= 3 else loop))
  polys.sort(key=lambda pv: min(pv[1]))
  return polys


def mass_properties(points: np.ndarray, tris: np.ndarray):
  """Volume, centre of mass and inertia tensor about it (unit density) of the closed triangle surface (outward orientation)."""
  pts = np.asarray(points, dtype=np.float64)
  vol, com, C = 0.0, np.zeros(3), np.zeros((3, 3))
  canon = np.full((3, 3), 1.0 / 120.0) + np.eye(3) / 120.0  # integral of x x^T over the unit tetrahedron
  for tri in tris:
    A = pts[tri].T  # columns a, b, c: tetrahedron (0, a, b, c)
    det = np.linalg.det(A)
    vol += det / 6.0
    com += det / 24.0 * A.sum(axis=1)
    C += det * A @ canon @ A.T
  com /= vol
  C -= vol * np.outer(com, com)  # second-moment matrix about the centre of mass
  inertia = np.trace(C) * np.eye(3) - C
  return vol, com, inertia


def process(vertices, faces=None, scale=(1.0, 1.0, 1.0)):
  """All tables of one mesh asset.  `faces` (n, 3) are optional: without them the hull triangles stand in (MuJoCo does the same
  for vertex-only meshes)."""
  from .mjcf import _principal, quat_to_mat

  v = np.asarray(vertices, dtype=np.float64).reshape(-1, 3) * np.asarray(scale, dtype=np.float64)
  hull_ids, tris = convex_hull(v)
  f = np.asarray(faces, dtype=np.int32).reshape(-1, 3) if faces is not None and len(faces) else tris
  vol, com, inertia = mass_properties(v, f if faces is not None and len(faces) else tris)
  diag, quat = _principal(inertia)
  R = quat_to_mat(quat)
  local = (v - com) @ R  # mesh frame: x_local = R^T (x - com)
  hull_ids, tris = convex_hull(local)
  polys = hull_polygons(local, tris)
  nvert = len(local)
  polymap = [[] for _ in range(nvert)]
  for p, (_, loop) in enumerate(polys):
    for vid in loop:
      polymap[vid].append(p)
  return dict(
    vert=local, face=f.astype(np.int32), graph=hull_graph(hull_ids, tris), pos=com, quat=quat,
    volume=vol, inertia=diag,
    polynormal=np.array([n for n, _ in polys]), polyvertnum=np.array([len(l) for _, l in polys], dtype=np.int32),
    polyvert=np.array([vid for _, l in polys for vid in l], dtype=np.int32),
    polymapnum=np.array([len(x) for x in polymap], dtype=np.int32), polymap=np.array([p for x in polymap for p in x], dtype=np.int32),
    aabb_center=0.5 * (local.max(axis=0) + local.min(axis=0)), aabb_size=0.5 * (local.max(axis=0) - local.min(axis=0)),
    rbound=float(np.linalg.norm(local, axis=1).max()),
  )


def read_obj(path: str):
  """Vertices and triangles of a Wavefront OBJ file (`v` and `f` records; polygons are fanned, texture / normal indices dropped)."""
  verts, faces = [], []
  with open(path, "r", errors="replace") as f:
    for line in f:
      t = line.split()
      if not t:
        continue
      if t[0] == "v":
        verts.append([float(t[1]), float(t[2]), float(t[3])])
      elif t[0] == "f":
        idx = [int(x.split("/")[0]) for x in t[1:]]
        idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
        for k in range(1, len(idx) - 1):
          faces.append([idx[0], idx[k], idx[k + 1]])
  return np.asarray(verts, dtype=np.float64).reshape(-1, 3), np.asarray(faces, dtype=np.int32).reshape(-1, 3)


def read_stl(path: str):
  """Vertices and triangles of an STL file (binary or ASCII); coincident vertices are merged like MuJoCo's loader does."""
  raw = open(path, "rb").read()
  tri = None
  if len(raw) >= 84:
    n = int(np.frombuffer(raw[80:84], dtype="<u4")[0])
    if len(raw) == 84 + 50 * n:
      rec = np.frombuffer(raw[84:], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=n)
      tri = rec["v"].astype(np.float64)
  if tri is None:
    pts = [[float(x) for x in line.split()[1:4]] for line in raw.decode("ascii", errors="replace").splitlines() if line.strip().startswith("vertex")]
    tri = np.asarray(pts, dtype=np.float64).reshape(-1, 3, 3)
  verts, inv = np.unique(tri.reshape(-1, 3), axis=0, return_inverse=True)
  return verts, inv.reshape(-1, 3).astype(np.int32)


def read_file(path: str):
  ext = path.lower().rsplit(".", 1)[-1]
  if ext == "obj":
    return read_obj(path)
  if ext == "stl":
    return read_stl(path)
  raise NotImplementedError(f"mesh file format .{ext} is not supported (OBJ and STL are): {path}")
