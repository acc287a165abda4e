"""Minimal pure-Python stand-in for the `warp` API surface used by the reference's @wp.func-level code.

NVIDIA Warp and MuJoCo are not installable offline, so the reference cannot run as shipped.  Its device functions
(`@wp.func` in collision_primitive_core.py, math.py, the solver's cost functions ...) are however plain Python once
`wp.vec3`, `wp.mat33`, `wp.where`, ... exist.  This module provides those names with VALUE semantics (indexing a matrix
returns a copy, augmented assignment rebinds) and double-precision scalars, and `load_reference_module()` imports a
reference source file from /root/reference UNMODIFIED against it.  It is used only by the golden-vector generators under
tools/ (run in the build container, where /root/reference exists); nothing at test or run time imports it.
"""

import importlib.util
import math as _m
import sys
import types as _t


class Vec:
  __slots__ = ("v",)
  _n = 3

  def __init__(self, *a):
    n = self._n
    if len(a) == 0:
      self.v = [0.0] * n
    elif len(a) == 1 and isinstance(a[0], Vec):
      self.v = list(a[0].v)
    elif len(a) == 1 and isinstance(a[0], (list, tuple)):
      self.v = [float(x) for x in a[0]]
    elif len(a) == 1:
      self.v = [float(a[0])] * n
    else:
      assert len(a) == n, (len(a), n)
      self.v = [float(x) for x in a]

  def _new(self, vals):
    o = type(self).__new__(type(self))
    o.v = vals
    return o

  def __getitem__(self, i):
    return self.v[i]

  def __setitem__(self, i, x):
    self.v[i] = float(x)

  def __len__(self):
    return len(self.v)

  def __iter__(self):
    return iter(self.v)

  x = property(lambda s: s.v[0])
  y = property(lambda s: s.v[1])
  z = property(lambda s: s.v[2])
  w = property(lambda s: s.v[3])

  def __neg__(self):
    return self._new([-a for a in self.v])

  def __add__(self, o):
    return self._new([a + b for a, b in zip(self.v, o.v)])

  def __sub__(self, o):
    return self._new([a - b for a, b in zip(self.v, o.v)])

  def __mul__(self, s):
    return self._new([a * float(s) for a in self.v])

  __rmul__ = __mul__

  def __truediv__(self, s):
    return self._new([a / float(s) for a in self.v])

  def __repr__(self):
    return f"vec{len(self.v)}({self.v})"


class Mat:
  __slots__ = ("m",)
  _shape = (3, 3)

  def __init__(self, *a):
    r, c = self._shape
    if len(a) == 0:
      self.m = [[0.0] * c for _ in range(r)]
    elif len(a) == 1 and isinstance(a[0], Mat):
      self.m = [list(row) for row in a[0].m]
    elif len(a) == 1:
      self.m = [[float(a[0])] * c for _ in range(r)]
    else:
      assert len(a) == r * c, (len(a), r, c)
      self.m = [[float(a[i * c + j]) for j in range(c)] for i in range(r)]

  def _new(self, rows):
    o = type(self).__new__(type(self))
    o.m = rows
    return o

  def __getitem__(self, idx):
    if isinstance(idx, tuple):
      return self.m[idx[0]][idx[1]]
    return _vec_cls(len(self.m[idx]))(list(self.m[idx]))  # a COPY of the row

  def __setitem__(self, idx, x):
    if isinstance(idx, tuple):
      self.m[idx[0]][idx[1]] = float(x)
    else:
      self.m[idx] = [float(t) for t in x]

  def __matmul__(self, o):
    if isinstance(o, Vec):
      return _vec_cls(len(self.m))([sum(a * b for a, b in zip(row, o.v)) for row in self.m])
    cols = list(zip(*o.m))
    return _mat_cls(len(self.m), len(cols))._from_rows([[sum(a * b for a, b in zip(row, col)) for col in cols] for row in self.m])

  def __mul__(self, o):  # warp allows mat * vec as well as scalar scaling
    if isinstance(o, (Vec, Mat)):
      return self.__matmul__(o)
    return self._new([[a * float(o) for a in row] for row in self.m])

  def __rmul__(self, s):
    return self._new([[a * float(s) for a in row] for row in self.m])

  def __add__(self, o):
    return self._new([[a + b for a, b in zip(r1, r2)] for r1, r2 in zip(self.m, o.m)])

  def __sub__(self, o):
    return self._new([[a - b for a, b in zip(r1, r2)] for r1, r2 in zip(self.m, o.m)])

  def __neg__(self):
    return self._new([[-a for a in row] for row in self.m])

  @classmethod
  def _from_rows(cls, rows):
    o = cls.__new__(cls)
    o.m = [list(map(float, r)) for r in rows]
    return o

  def __repr__(self):
    return f"mat({self.m})"


_VEC, _MAT = {}, {}


def _vec_cls(n):
  if n not in _VEC:
    _VEC[n] = type(f"vec{n}", (Vec,), {"_n": n, "__slots__": ()})
  return _VEC[n]


def _mat_cls(r, c):
  if (r, c) not in _MAT:
    _MAT[(r, c)] = type(f"mat{r}{c}", (Mat,), {"_shape": (r, c), "__slots__": ()})
  return _MAT[(r, c)]


def _build_warp():
  wp = _t.ModuleType("warp")
  wp.func = lambda f: f
  wp.kernel = lambda f=None, **k: f if f is not None else (lambda g: g)
  wp.struct = lambda c: c
  wp.set_module_options = lambda *a, **k: None
  wp.static = lambda x: x
  wp.vec2, wp.vec3, wp.vec4 = _vec_cls(2), _vec_cls(3), _vec_cls(4)
  wp.vec2f, wp.vec3f, wp.vec4f, wp.quat, wp.quatf = wp.vec2, wp.vec3, wp.vec4, _vec_cls(4), _vec_cls(4)
  wp.vec2i = wp.vec2
  wp.spatial_vector = _vec_cls(6)
  wp.mat33, wp.mat33f = _mat_cls(3, 3), _mat_cls(3, 3)
  wp.float32 = wp.float64 = wp.float = float
  wp.int32 = wp.int = int
  wp.bool = bool
  wp.inf, wp.pi = _m.inf, _m.pi
  wp.array = wp.array2d = wp.array3d = wp.array4d = type("array", (), {"__class_getitem__": classmethod(lambda c, k: c)})
  tt = _t.ModuleType("warp.types")
  tt.vector = lambda length, dtype=float: _vec_cls(length)
  tt.matrix = lambda shape, dtype=float: _mat_cls(*shape)
  wp.types = tt

  def where(c, a, b):
    return a if c else b

  def _abs(x):
    return x._new([abs(a) for a in x.v]) if isinstance(x, Vec) else abs(x)

  def _ew(f):
    def g(a, b):
      if isinstance(a, Vec):
        return a._new([f(p, q) for p, q in zip(a.v, b.v)])
      return f(a, b)
    return g

  wp.where = where
  wp.abs = _abs
  wp.min, wp.max = _ew(min), _ew(max)
  wp.dot = lambda a, b: sum(p * q for p, q in zip(a.v, b.v))
  wp.cross = lambda a, b: a._new([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])
  wp.length_sq = lambda a: sum(p * p for p in a.v)
  wp.length = lambda a: _m.sqrt(sum(p * p for p in a.v))
  wp.norm_l2 = wp.length

  def normalize(a):
    l = wp.length(a)
    return a / l if l > 0.0 else a._new([0.0] * len(a.v))

  wp.normalize = normalize
  wp.cw_mul = lambda a, b: a._new([p * q for p, q in zip(a.v, b.v)])
  wp.sqrt, wp.sin, wp.cos, wp.atan2, wp.acos, wp.exp, wp.log, wp.pow = _m.sqrt, _m.sin, _m.cos, _m.atan2, _m.acos, _m.exp, _m.log, _m.pow
  wp.sign = lambda x: -1.0 if x < 0 else 1.0  # warp: sign(0) = +1
  wp.clamp = lambda x, lo, hi: min(max(x, lo), hi)
  wp.transpose = lambda a: _mat_cls(len(a.m[0]), len(a.m))._from_rows([list(c) for c in zip(*a.m)])
  wp.matrix_from_rows = lambda *rows: _mat_cls(len(rows), len(rows[0].v))._from_rows([list(r.v) for r in rows])
  wp.matrix_from_cols = lambda *cols: wp.transpose(wp.matrix_from_rows(*cols))
  wp.identity = lambda n, dtype=float: _mat_cls(n, n)._from_rows([[1.0 if i == j else 0.0 for j in range(n)] for i in range(n)])
  return wp


def install(extra_types=None):
  """Put the shim `warp` and a constants-only `mujoco_warp._src.types` into sys.modules; returns the fake warp module."""
  wp = _build_warp()
  sys.modules["warp"] = wp
  sys.modules["warp.types"] = wp.types
  pkg = _t.ModuleType("mujoco_warp")
  pkg.__path__ = []
  sub = _t.ModuleType("mujoco_warp._src")
  sub.__path__ = []
  ty = _t.ModuleType("mujoco_warp._src.types")
  # numeric constants restated from /root/reference/mujoco_warp/_src/types.py:32-56 (MJ_MINVAL = mujoco.mjMINVAL = 1e-15)
  ty.MJ_MINVAL, ty.MJ_MAXVAL, ty.MJ_MINIMP, ty.MJ_MAXIMP, ty.MJ_MINMU = 1e-15, 1e10, 0.0001, 0.9999, 1e-5
  ty.vec5, ty.vec6, ty.vec8, ty.vec10, ty.vec11 = (_vec_cls(n) for n in (5, 6, 8, 10, 11))
  ty.vec10f = ty.vec10
  for k, v in (extra_types or {}).items():
    setattr(ty, k, v)
  sys.modules["mujoco_warp"], sys.modules["mujoco_warp._src"], sys.modules["mujoco_warp._src.types"] = pkg, sub, ty
  sub.types = ty
  return wp


def load_reference_module(name, root="/root/reference/mujoco_warp/_src"):
  """Import /root/reference/mujoco_warp/_src/<name>.py unmodified under the shim (install() must have been called)."""
  full = f"mujoco_warp._src.{name}"
  spec = importlib.util.spec_from_file_location(full, f"{root}/{name}.py")
  mod = importlib.util.module_from_spec(spec)
  sys.modules[full] = mod
  spec.loader.exec_module(mod)
  setattr(sys.modules["mujoco_warp._src"], name, mod)
  return mod
