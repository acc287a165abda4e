"""ctypes wrapper of the CPU oracle (oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this
module.  The product package `mujoco_warp_b200` never does.

The derived index tables below (filtered NXN geom pairs, body_isdofancestor, limited-joint list, qLD block
offsets) restate /root/reference/mujoco_warp/_src/io.py:495-640,536-549 independently of the
product's `io.put_model`, so that a bug there cannot hide in both sides of a parity test.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(force: bool = False):
  """Compile oracle.c into liborc64.so (double) and liborc32.so (float) next to this file."""
  src = os.path.join(_HERE, "oracle.c")
  for name, flags in (("liborc64.so", []), ("liborc32.so", ["-DORC_FLOAT"])):
    out = os.path.join(_HERE, name)
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "oracle.h")), os.path.getmtime(os.path.join(_HERE, "oracle_ccd.h"))):
      continue
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-std=c99", "-fno-fast-math", "-ffp-contract=off", *flags, src, "-o", out, "-lm"]
    subprocess.check_call(cmd)


def _lib(real_bytes: int):
  if real_bytes not in _LIBS:
    build()
    lib = ctypes.CDLL(os.path.join(_HERE, "liborc64.so" if real_bytes == 8 else "liborc32.so"))
    lib.orc_model_create.restype = ctypes.c_void_p
    lib.orc_data_create.restype = ctypes.c_void_p
    lib.orc_data_create.argtypes = [ctypes.c_int] * 3
    lib.orc_last_error.restype = ctypes.c_char_p
    lib.orc_halton.restype = ctypes.c_double
    lib.orc_halton.argtypes = [ctypes.c_int, ctypes.c_int]
    for f in ("orc_model_set_int",):
      getattr(lib, f).argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
    lib.orc_model_set_real.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double]
    for f in ("orc_model_set_iarr", "orc_model_set_rarr", "orc_data_set_iarr", "orc_data_set_rarr"):
      getattr(lib, f).argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
    for f in ("orc_forward", "orc_step"):
      getattr(lib, f).argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    lib.orc_model_free.argtypes = [ctypes.c_void_p]
    lib.orc_data_free.argtypes = [ctypes.c_void_p]
    assert lib.orc_sizeof_real() == real_bytes
    _LIBS[real_bytes] = lib
  return _LIBS[real_bytes]


MODEL_INTS = ["nq", "nv", "nu", "nbody", "nmocap", "njnt", "ngeom", "nsite", "ncam", "nlight", "nC", "ntree"]
MODEL_IARRS = [
  "body_parentid", "body_rootid", "body_weldid", "body_mocapid", "body_jntnum", "body_jntadr", "body_dofnum", "body_dofadr",
  "jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_bodyid", "jnt_actfrclimited", "jnt_actgravcomp",
  "dof_bodyid", "dof_jntid", "dof_parentid", "M_rownnz", "M_rowadr", "M_colind", "tree_dofadr", "tree_dofnum",
  "geom_type", "geom_condim", "geom_bodyid", "geom_priority",
  "actuator_trnid", "actuator_gaintype", "actuator_biastype", "actuator_ctrllimited", "actuator_forcelimited",
  "cam_mode", "cam_bodyid", "cam_targetbodyid", "light_mode", "light_bodyid", "light_targetbodyid", "site_bodyid",
]
MODEL_RARRS = [
  "qpos0", "qpos_spring", "body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_subtreemass",
  "body_inertia", "body_invweight0", "body_gravcomp", "jnt_pos", "jnt_axis", "jnt_stiffness", "jnt_range", "jnt_margin", "jnt_solref",
  "jnt_solimp", "jnt_actfrcrange", "dof_armature", "dof_damping", "dof_invweight0", "dof_frictionloss", "dof_solref",
  "dof_solimp", "geom_size", "geom_aabb", "geom_rbound", "geom_pos", "geom_quat", "geom_friction", "geom_margin",
  "geom_gap", "geom_solmix", "geom_solref", "geom_solimp", "actuator_gear", "actuator_gainprm", "actuator_biasprm",
  "actuator_ctrlrange", "actuator_forcerange", "cam_pos", "cam_quat", "cam_poscom0", "cam_pos0", "cam_mat0",
  "light_pos", "light_dir", "light_poscom0", "light_pos0", "light_dir0", "site_pos", "site_quat",
]


def derived_tables(mjm):
  """Restates io.py:536-640 (isdofancestor, filtered NXN pairs) and io.py:748- (limited joints)."""
  nbody, nv, ngeom = mjm.nbody, mjm.nv, mjm.ngeom
  anc = np.zeros((nbody, nv), dtype=np.int32)
  for bodyid in range(nbody):
    b = bodyid
    while b > 0 and mjm.body_dofnum[b] == 0:
      b = mjm.body_parentid[b]
    if mjm.body_dofnum[b] == 0:
      continue
    dofid = mjm.body_dofadr[b] + mjm.body_dofnum[b] - 1
    while dofid >= 0:
      anc[bodyid, dofid] = 1
      dofid = mjm.dof_parentid[dofid]
  filterparent = not (mjm.opt.disableflags & (1 << 10))
  g1, g2 = np.triu_indices(ngeom, k=1)
  b1, b2 = mjm.geom_bodyid[g1], mjm.geom_bodyid[g2]
  w1, w2 = mjm.body_weldid[b1], mjm.body_weldid[b2]
  wp1, wp2 = mjm.body_weldid[mjm.body_parentid[w1]], mjm.body_weldid[mjm.body_parentid[w2]]
  self_col = w1 == w2
  parent_child = filterparent & (w1 != 0) & (w2 != 0) & ((w1 == wp2) | (w2 == wp1))
  mask = ((mjm.geom_contype[g1] & mjm.geom_conaffinity[g2]) | (mjm.geom_contype[g2] & mjm.geom_conaffinity[g1])).astype(bool)
  exclude = np.isin((b1.astype(np.int64) << 16) + b2, mjm.exclude_signature)
  pairid = -np.ones(len(g1), dtype=np.int32)
  pairid[~(mask & ~self_col & ~parent_child & ~exclude)] = -2
  for i in range(int(getattr(mjm, "npair", 0))):  # explicit pairs override the filters (io.py:577-583)
    a, b = sorted((int(mjm.pair_geom1[i]), int(mjm.pair_geom2[i])))
    pairid[(a * (2 * ngeom - a - 3)) // 2 + b - 1] = i
  include = pairid > -2
  pairs = np.stack((g1, g2), axis=1)[include].astype(np.int32)
  pid = np.stack((pairid[include], -np.ones(include.sum(), dtype=np.int32)), axis=1).astype(np.int32)
  limited = np.nonzero(np.asarray(mjm.jnt_limited).astype(bool) & ((mjm.jnt_type == 2) | (mjm.jnt_type == 3)))[0].astype(np.int32)
  limited_ball = np.nonzero(np.asarray(mjm.jnt_limited).astype(bool) & (mjm.jnt_type == 1))[0].astype(np.int32)
  blk = np.zeros(max(nv, 1), dtype=np.int32)
  off = 0
  for adr, num in zip(mjm.tree_dofadr, mjm.tree_dofnum):
    blk[adr : adr + num] = off
    off += int(num) * int(num)
  nJmom = 0
  for i in range(mjm.nu):
    if int(np.asarray(getattr(mjm, "actuator_trntype", np.zeros(mjm.nu)))[i]) == 3:  # tendon transmission: the tendon's Jacobian row
      nJmom += int(mjm.ten_J_rownnz[mjm.actuator_trnid[i, 0]])
      continue
    t = mjm.jnt_type[mjm.actuator_trnid[i, 0]]
    nJmom += {0: 6, 1: 3, 2: 1, 3: 1}[int(t)]
  nmaxcondim = int(mjm.geom_condim.max()) if ngeom else 1
  if getattr(mjm, "npair", 0):
    nmaxcondim = max(nmaxcondim, int(np.asarray(mjm.pair_dim).max()))
  # collision_convex.py:1209-1223: EPA gets 16 iterations when every convex pair of the model is box-box
  convex = {(2, 4), (3, 4), (3, 5), (4, 4), (4, 5), (4, 6), (5, 5), (5, 6), (2, 7), (3, 7), (4, 7), (5, 7), (6, 7), (7, 7)}
  if not (int(mjm.opt.disableflags) & (1 << 17)):
    convex.add((6, 6))
  gt = np.asarray(mjm.geom_type)
  keys = [(min(int(gt[a]), int(gt[b])), max(int(gt[a]), int(gt[b]))) for a, b in pairs]
  nconvex = sum(k in convex for k in keys)
  nboxbox = sum(k == (6, 6) and k in convex for k in keys)
  epa_iterations = 16 if nboxbox == nconvex else int(getattr(mjm.opt, "ccd_iterations", 35))
  return dict(epa_iterations=epa_iterations, body_isdofancestor=anc, nxn_geom_pair=pairs, nxn_pairid=pid, jnt_limited_slide_hinge_adr=limited, jnt_limited_ball_adr=limited_ball,
              qLD_block_adr=blk, qld_total=off, nJmom=nJmom, nmaxpyramid=max(1, 2 * (nmaxcondim - 1)))


def data_spec(mjm, tabs, nconmax, njmax):
  """name -> (is_int, per-world shape)."""
  nb, nv, nq, nu, nj, ng = mjm.nbody, mjm.nv, mjm.nq, mjm.nu, mjm.njnt, mjm.ngeom
  npyr = tabs["nmaxpyramid"]
  R, I = False, True
  return {
    "time": (R, ()), "qpos": (R, (nq,)), "qvel": (R, (nv,)), "ctrl": (R, (nu,)), "qacc_warmstart": (R, (nv,)),
    "act": (R, (int(getattr(mjm, "na", 0)),)), "act_dot": (R, (int(getattr(mjm, "na", 0)),)),
    "ten_length": (R, (int(getattr(mjm, "ntendon", 0)),)), "ten_velocity": (R, (int(getattr(mjm, "ntendon", 0)),)),
    "ten_J": (R, (int(getattr(mjm, "nJten", 0)) if int(getattr(mjm, "ntendon", 0)) else 0,)),
    "qfrc_applied": (R, (nv,)), "xfrc_applied": (R, (nb, 6)), "qacc": (R, (nv,)),
    "mocap_pos": (R, (int(getattr(mjm, "nmocap", 0)), 3)), "mocap_quat": (R, (int(getattr(mjm, "nmocap", 0)), 4)),
    "xpos": (R, (nb, 3)), "xquat": (R, (nb, 4)), "xmat": (R, (nb, 3, 3)), "xipos": (R, (nb, 3)), "ximat": (R, (nb, 3, 3)),
    "xanchor": (R, (nj, 3)), "xaxis": (R, (nj, 3)), "geom_xpos": (R, (ng, 3)), "geom_xmat": (R, (ng, 3, 3)),
    "site_xpos": (R, (mjm.nsite, 3)), "site_xmat": (R, (mjm.nsite, 3, 3)),
    "cam_xpos": (R, (mjm.ncam, 3)), "cam_xmat": (R, (mjm.ncam, 3, 3)), "light_xpos": (R, (mjm.nlight, 3)), "light_xdir": (R, (mjm.nlight, 3)),
    "subtree_com": (R, (nb, 3)), "cdof": (R, (nv, 6)), "cinert": (R, (nb, 10)), "crb": (R, (nb, 10)), "M": (R, (mjm.nC,)),
    "qLD": (R, (tabs["qld_total"],)),
    "actuator_length": (R, (nu,)), "actuator_moment": (R, (tabs["nJmom"],)), "actuator_velocity": (R, (nu,)),
    "cvel": (R, (nb, 6)), "cdof_dot": (R, (nv, 6)), "qfrc_bias": (R, (nv,)), "qfrc_spring": (R, (nv,)), "qfrc_damper": (R, (nv,)),
    "qfrc_gravcomp": (R, (nv,)), "qfrc_passive": (R, (nv,)), "actuator_force": (R, (nu,)), "qfrc_actuator": (R, (nv,)),
    "qfrc_smooth": (R, (nv,)), "qacc_smooth": (R, (nv,)), "qfrc_constraint": (R, (nv,)), "cacc": (R, (nb, 6)), "cfrc_int": (R, (nb, 6)), "cfrc_ext": (R, (nb, 6)),
    "sensordata": (R, (int(getattr(mjm, "nsensordata", 0)) if int(getattr(mjm, "nsensor", 0)) else 0,)), "subtree_linvel": (R, (nb, 3)), "subtree_angmom": (R, (nb, 3)),
    "efc_J": (R, (njmax, nv)), "efc_pos": (R, (njmax,)), "efc_margin": (R, (njmax,)), "efc_D": (R, (njmax,)), "efc_vel": (R, (njmax,)),
    "efc_aref": (R, (njmax,)), "efc_frictionloss": (R, (njmax,)), "efc_force": (R, (njmax,)), "efc_Ma": (R, (nv,)),
    "con_dist": (R, (nconmax,)), "con_pos": (R, (nconmax, 3)), "con_frame": (R, (nconmax, 3, 3)), "con_includemargin": (R, (nconmax,)),
    "con_friction": (R, (nconmax, 5)), "con_solref": (R, (nconmax, 2)), "con_solreffriction": (R, (nconmax, 2)), "con_solimp": (R, (nconmax, 5)),
    "ne": (I, ()), "nf": (I, ()), "nl": (I, ()), "nefc": (I, ()), "ncon": (I, ()), "ncollision": (I, ()), "solver_niter": (I, ()), "overflow": (I, ()),
    "efc_type": (I, (njmax,)), "efc_id": (I, (njmax,)), "efc_state": (I, (njmax,)),
    "moment_rownnz": (I, (nu,)), "moment_rowadr": (I, (nu,)), "moment_colind": (I, (tabs["nJmom"],)),
    "eq_active": (I, (int(getattr(mjm, "neq", 0)),)),
    "con_dim": (I, (nconmax,)), "con_geom": (I, (nconmax, 2)), "con_efc_address": (I, (nconmax, npyr)), "con_geomcollisionid": (I, (nconmax,)),
  }


class Oracle:
  """Batched CPU oracle bound to one compiled model.

  o = Oracle(mjm, nworld=4, nconmax=24, njmax=64); o.d["qpos"][:] = ...; o.forward(); o.d["qacc"]
  """

  def __init__(self, mjm, nworld=1, nconmax=24, njmax=64, dtype=np.float64, static_kin=None, clamp_tolerance=True):
    self.real = np.dtype(dtype)
    self.lib = _lib(self.real.itemsize)
    self.mjm, self.nworld, self.nconmax, self.njmax = mjm, nworld, nconmax, njmax
    self.tabs = derived_tables(mjm)
    lib = self.lib
    self._keep = []
    self.m = ctypes.c_void_p(lib.orc_model_create())

    def seti(name, v):
      assert lib.orc_model_set_int(self.m, name.encode(), int(v)) == 0, name

    def setr(name, v):
      assert lib.orc_model_set_real(self.m, name.encode(), float(v)) == 0, name

    def setia(name, arr):
      a = np.ascontiguousarray(np.asarray(arr).astype(np.int32))
      self._keep.append(a)
      assert lib.orc_model_set_iarr(self.m, name.encode(), a.ctypes.data) == 0, name

    def setra(name, arr):
      a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64).astype(self.real))
      self._keep.append(a)
      assert lib.orc_model_set_rarr(self.m, name.encode(), a.ctypes.data) == 0, name

    for n in MODEL_INTS:
      seti(n, getattr(mjm, n))
    o = mjm.opt
    seti("nJmom", self.tabs["nJmom"]); seti("nxn_npair", len(self.tabs["nxn_geom_pair"])); seti("nlimit", len(self.tabs["jnt_limited_slide_hinge_adr"]))
    seti("nmaxpyramid", self.tabs["nmaxpyramid"])
    neq = int(getattr(mjm, "neq", 0))
    seti("nlimit_ball", len(self.tabs["jnt_limited_ball_adr"])); seti("neq", neq)
    seti("integrator", o.integrator); seti("cone", o.cone); seti("solver", o.solver); seti("iterations", o.iterations)
    seti("ls_iterations", o.ls_iterations); seti("disableflags", o.disableflags); seti("enableflags", o.enableflags)
    seti("broadphase_filter", getattr(o, "broadphase_filter", 1 | 2 | 8))  # io.py:405 default PLANE|SPHERE|OBB
    seti("broadphase", getattr(o, "broadphase", 0))  # 0 NXN (io.py:404), 1 / 2 sweep-and-prune
    seti("ccd_iterations", getattr(o, "ccd_iterations", 35)); setr("ccd_tolerance", getattr(o, "ccd_tolerance", 1e-6))
    seti("epa_iterations", self.tabs["epa_iterations"])
    tol = float(o.tolerance)
    if clamp_tolerance:
      tol = max(tol, 1e-6)  # io.py:401: put_model clamps the solver tolerance (chosen for float32) whatever the host precision;
      # clamp_tolerance=False lets invariant tests converge the fp64 build fully
    self.tolerance = tol
    setr("timestep", o.timestep); setr("tolerance", tol); setr("ls_tolerance", o.ls_tolerance)
    setr("impratio_invsqrt", 1.0 / np.sqrt(o.impratio)); setr("meaninertia", mjm.stat.meaninertia)
    for n in MODEL_IARRS:
      setia(n, getattr(mjm, n))
    for n in MODEL_RARRS:
      setra(n, getattr(mjm, n))
    setra("gravity", o.gravity)
    for n in ("nxn_geom_pair", "nxn_pairid", "jnt_limited_slide_hinge_adr", "jnt_limited_ball_adr", "body_isdofancestor", "qLD_block_adr"):
      setia(n, self.tabs[n])
    for n in ("eq_type", "eq_obj1id", "eq_obj2id"):
      setia(n, getattr(mjm, n) if neq else np.zeros(1, dtype=np.int32))
    npair = int(getattr(mjm, "npair", 0))
    setia("pair_dim", mjm.pair_dim if npair else np.zeros(1, dtype=np.int32))
    for n, k in (("pair_friction", 5), ("pair_solref", 2), ("pair_solreffriction", 2), ("pair_solimp", 5), ("pair_margin", 1), ("pair_gap", 1)):
      setra(n, getattr(mjm, n) if npair else np.zeros(k))
    for n, k in (("eq_solref", 2), ("eq_solimp", 5), ("eq_data", 11)):
      setra(n, getattr(mjm, n) if neq else np.zeros(k))
    nsensor = int(getattr(mjm, "nsensor", 0))
    seti("nsensor", nsensor); seti("nsensordata", int(getattr(mjm, "nsensordata", 0)) if nsensor else 0)
    for n in ("sensor_type", "sensor_datatype", "sensor_needstage", "sensor_objtype", "sensor_objid", "sensor_reftype", "sensor_refid", "sensor_dim", "sensor_adr"):
      setia(n, getattr(mjm, n) if nsensor else np.zeros(1, dtype=np.int32))
    setra("sensor_cutoff", mjm.sensor_cutoff if nsensor else np.zeros(1))
    nmesh = int(getattr(mjm, "nmesh", 0))
    seti("nmesh", nmesh)
    setia("geom_dataid", getattr(mjm, "geom_dataid", -np.ones(mjm.ngeom, dtype=np.int32)))
    for n in ("mesh_vertadr", "mesh_vertnum", "mesh_graphadr", "mesh_graph", "mesh_polynum", "mesh_polyadr", "mesh_polyvertadr", "mesh_polyvertnum",
              "mesh_polyvert", "mesh_polymapadr", "mesh_polymapnum", "mesh_polymap"):
      setia(n, getattr(mjm, n) if nmesh and len(np.asarray(getattr(mjm, n))) else np.zeros(1, dtype=np.int32))
    setra("mesh_vert", mjm.mesh_vert if nmesh else np.zeros(3)); setra("mesh_polynormal", mjm.mesh_polynormal if nmesh else np.zeros(3))
    na, nu = int(getattr(mjm, "na", 0)), int(mjm.nu)
    seti("na", na)
    adr_default = -np.ones(max(nu, 1), dtype=np.int32)
    for n, dflt in (("actuator_dyntype", np.zeros(max(nu, 1))), ("actuator_actadr", adr_default), ("actuator_actnum", np.zeros(max(nu, 1))),
                    ("actuator_actlimited", np.zeros(max(nu, 1))), ("actuator_actearly", np.zeros(max(nu, 1)))):
      setia(n, getattr(mjm, n, dflt) if nu else dflt)
    setra("actuator_dynprm", getattr(mjm, "actuator_dynprm", np.zeros((max(nu, 1), 10))) if nu else np.zeros(10))
    setra("actuator_actrange", getattr(mjm, "actuator_actrange", np.zeros((max(nu, 1), 2))) if nu else np.zeros(2))
    nt = int(getattr(mjm, "ntendon", 0))
    seti("ntendon", nt); seti("nJten", int(getattr(mjm, "nJten", 0)) if nt else 0)
    setia("actuator_trntype", getattr(mjm, "actuator_trntype", np.zeros(max(nu, 1))) if nu else np.zeros(1))
    for n in ("ten_J_rownnz", "ten_J_rowadr", "ten_J_colind", "tendon_adr", "tendon_num", "wrap_objid", "tendon_limited", "tendon_actfrclimited"):
      setia(n, np.asarray(getattr(mjm, n)).astype(np.int32) if nt else np.zeros(1, dtype=np.int32))
    for n, k in (("wrap_prm", 1), ("tendon_range", 2), ("tendon_margin", 1), ("tendon_stiffness", 1), ("tendon_damping", 1), ("tendon_frictionloss", 1),
                 ("tendon_lengthspring", 2), ("tendon_length0", 1), ("tendon_invweight0", 1), ("tendon_solref_lim", 2), ("tendon_solimp_lim", 5),
                 ("tendon_solref_fri", 2), ("tendon_solimp_fri", 5), ("tendon_actfrcrange", 2)):
      setra(n, getattr(mjm, n) if nt else np.zeros(k))
    if nt and np.any(np.asarray(getattr(mjm, "tendon_armature", 0)) != 0):
      raise NotImplementedError("oracle: tendon armature is not restated")
    nsite = int(getattr(mjm, "nsite", 0))
    setia("site_type", getattr(mjm, "site_type", 2 * np.ones(nsite, dtype=np.int32)) if nsite else np.zeros(1, dtype=np.int32))
    setra("site_size", getattr(mjm, "site_size", 0.005 * np.ones((nsite, 3))) if nsite else np.zeros(3))

    self.spec = data_spec(mjm, self.tabs, nconmax, njmax)
    self.dptr = ctypes.c_void_p(lib.orc_data_create(nworld, nconmax, njmax))
    self.d = {}
    for name, (is_int, shape) in self.spec.items():
      a = np.zeros((nworld,) + tuple(shape), dtype=np.int32 if is_int else self.real)
      self.d[name] = a
      fn = lib.orc_data_set_iarr if is_int else lib.orc_data_set_rarr
      assert fn(self.dptr, name.encode(), a.ctypes.data) == 0, name
    # initial state: qpos0, static geom poses from host kinematics (io.py:1815-1843)
    self.d["qpos"][:] = np.asarray(mjm.qpos0)
    if neq:
      self.d["eq_active"][:] = np.asarray(mjm.eq_active0).astype(np.int32)
    if getattr(mjm, "nmocap", 0):  # io.py:1824-1846: mocap poses start at the bodies' model pose
      mb = np.nonzero(np.asarray(mjm.body_mocapid) >= 0)[0]
      order = mb[np.argsort(np.asarray(mjm.body_mocapid)[mb])]
      self.d["mocap_pos"][:] = np.asarray(mjm.body_pos)[order]
      self.d["mocap_quat"][:] = np.asarray(mjm.body_quat)[order]
    if static_kin is not None:
      self.d["geom_xpos"][:] = static_kin.geom_xpos
      self.d["geom_xmat"][:] = static_kin.geom_xmat

  def set_state(self, qpos=None, qvel=None, ctrl=None, qacc_warmstart=None, time=None, act=None):
    for k, v in (("qpos", qpos), ("qvel", qvel), ("ctrl", ctrl), ("qacc_warmstart", qacc_warmstart), ("time", time), ("act", act)):
      if v is not None:
        self.d[k][...] = np.asarray(v)

  def forward(self, nthreads=0):
    if self.lib.orc_forward(self.m, self.dptr, nthreads) != 0:
      raise RuntimeError(self.lib.orc_last_error().decode())

  def step(self, nthreads=0):
    if self.lib.orc_step(self.m, self.dptr, nthreads) != 0:
      raise RuntimeError(self.lib.orc_last_error().decode())

  def __del__(self):
    try:
      self.lib.orc_model_free(self.m)
      self.lib.orc_data_free(self.dptr)
    except Exception:
      pass


def halton(index, base):
  return _lib(8).orc_halton(int(index), int(base))


def ccd(type1, size1, pos1, mat1, type2, size2, pos2, mat2, margin=0.0, tolerance=1e-6, cutoff=1e30, iterations=35, multiccd=False, dtype=np.float64):
  """Convex pair routine on two posed geoms -> (dist, ncon, witness1[4,3], witness2[4,3], overflow).
  Mirrors the harness of the reference's GJK tests (collision_gjk_test.py:35-303 `_geom_dist`)."""
  real = np.dtype(dtype)
  lib = _lib(real.itemsize)
  c_real = ctypes.c_double if real.itemsize == 8 else ctypes.c_float
  arr = lambda a, n: np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1)[:n].astype(real))
  s1, p1, m1, s2, p2, m2 = arr(size1, 3), arr(pos1, 3), arr(mat1, 9), arr(size2, 3), arr(pos2, 3), arr(mat2, 9)
  dist = np.zeros(1, dtype=real); w1 = np.zeros((4, 3), dtype=real); w2 = np.zeros((4, 3), dtype=real); ovf = np.zeros(1, dtype=np.int32)
  P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
  lib.orc_ccd.restype = ctypes.c_int
  lib.orc_ccd.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                          c_real, c_real, c_real, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
  n = lib.orc_ccd(int(type1), P(s1), P(p1), P(m1), int(type2), P(s2), P(p2), P(m2), margin, tolerance, cutoff, int(iterations), int(bool(multiccd)),
                  P(dist), P(w1), P(w2), P(ovf))
  return float(dist[0]), int(n), w1, w2, int(ovf[0])
