"""Runs the reference's own host pipeline (io.put_model / make_data / forward / step and every kernel they launch, sources
UNMODIFIED from /root/reference/mujoco_warp/_src) on the CPU through tools/warp_shim.py, starting from a model compiled by
mujoco_warp_b200._src.mjcf.  Neither NVIDIA Warp nor MuJoCo is installable offline; the shim emulates warp's kernel
launch semantics in Python and `fake_mujoco()` supplies the handful of MuJoCo enums/constants/functions io.py touches.

Only the golden generators under tools/ use this (build container only); tests read the committed fixtures.
"""

import enum
import os
import sys
import types as _t

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import warp_shim  # noqa: E402

_ENUMS = {
  "mjtDisableBit": dict(CONSTRAINT=1 << 0, EQUALITY=1 << 1, FRICTIONLOSS=1 << 2, LIMIT=1 << 3, CONTACT=1 << 4, SPRING=1 << 5, DAMPER=1 << 6,
                        GRAVITY=1 << 7, CLAMPCTRL=1 << 8, WARMSTART=1 << 9, FILTERPARENT=1 << 10, ACTUATION=1 << 11, REFSAFE=1 << 12,
                        SENSOR=1 << 13, MIDPHASE=1 << 14, EULERDAMP=1 << 15, AUTORESET=1 << 16, NATIVECCD=1 << 17, ISLAND=1 << 18, MULTICCD=1 << 19),
  "mjtEnableBit": dict(OVERRIDE=1, ENERGY=2, FWDINV=4, INVDISCRETE=8, MULTICCD=16, SLEEP=32),
  "mjtJoint": dict(FREE=0, BALL=1, SLIDE=2, HINGE=3),
  "mjtGeom": dict(PLANE=0, HFIELD=1, SPHERE=2, CAPSULE=3, ELLIPSOID=4, CYLINDER=5, BOX=6, MESH=7, SDF=8, FLEX=107),
  "mjtObj": dict(UNKNOWN=0, BODY=1, XBODY=2, JOINT=3, DOF=4, GEOM=5, SITE=6, CAMERA=7, LIGHT=8, FLEX=9, MESH=10, SKIN=11, HFIELD=12, TEXTURE=13,
                 MATERIAL=14, PAIR=15, EXCLUDE=16, EQUALITY=17, TENDON=18, ACTUATOR=19, SENSOR=20, NUMERIC=21, TEXT=22, TUPLE=23, KEY=24, PLUGIN=25),
  "mjtTrn": dict(JOINT=0, JOINTINPARENT=1, SLIDERCRANK=2, TENDON=3, SITE=4, BODY=5),
  "mjtDyn": dict(NONE=0, INTEGRATOR=1, FILTER=2, FILTEREXACT=3, MUSCLE=4, DCMOTOR=5, USER=6),
  "mjtGain": dict(FIXED=0, AFFINE=1, MUSCLE=2, DCMOTOR=3, USER=4),
  "mjtBias": dict(NONE=0, AFFINE=1, MUSCLE=2, DCMOTOR=3, USER=4),
  "mjtEq": dict(CONNECT=0, WELD=1, JOINT=2, TENDON=3, FLEX=4, FLEXSTRAIN=5),
  "mjtWrap": dict(NONE=0, JOINT=1, PULLEY=2, SITE=3, SPHERE=4, CYLINDER=5),
  "mjtIntegrator": dict(EULER=0, RK4=1, IMPLICIT=2, IMPLICITFAST=3),
  "mjtCone": dict(PYRAMIDAL=0, ELLIPTIC=1),
  "mjtSolver": dict(PGS=0, CG=1, NEWTON=2),
  "mjtJacobian": dict(DENSE=0, SPARSE=1, AUTO=2),
  "mjtConstraint": dict(EQUALITY=0, FRICTION_DOF=1, FRICTION_TENDON=2, LIMIT_JOINT=3, LIMIT_TENDON=4, CONTACT_FRICTIONLESS=5, CONTACT_PYRAMIDAL=6,
                        CONTACT_ELLIPTIC=7),
  "mjtConstraintState": dict(SATISFIED=0, QUADRATIC=1, LINEARNEG=2, LINEARPOS=3, CONE=4),
  "mjtCamLight": dict(FIXED=0, TRACK=1, TRACKCOM=2, TARGETBODY=3, TARGETBODYCOM=4),
  "mjtStage": dict(NONE=0, POS=1, VEL=2, ACC=3),
  "mjtDataType": dict(REAL=0, POSITIVE=1, AXIS=2, QUATERNION=3),
  "mjtProjection": dict(PERSPECTIVE=0, ORTHOGRAPHIC=1),
  "mjtSleepPolicy": dict(AUTO=0, AUTO_NEVER=1, AUTO_ALLOWED=2, NEVER=3, ALLOWED=4, INIT=5),
  "mjtSleepState": dict(STATIC=-1, ASLEEP=0, AWAKE=1),
  "mjtTextureRole": dict(USER=0, RGB=1),
  "mjtTexture": dict(TWO_D=0, CUBE=1, SKYBOX=2),
  "mjtLightType": dict(SPOT=0, DIRECTIONAL=1, POINT=2, IMAGE=3),
  "mjtCamOutBit": dict(RGB=1, DEPTH=2, DIST=4, NORMAL=8, SEG=16),
}
_PREFIX = {
  "mjtDisableBit": "mjDSBL_", "mjtEnableBit": "mjENBL_", "mjtJoint": "mjJNT_", "mjtGeom": "mjGEOM_", "mjtObj": "mjOBJ_", "mjtTrn": "mjTRN_",
  "mjtDyn": "mjDYN_", "mjtGain": "mjGAIN_", "mjtBias": "mjBIAS_", "mjtEq": "mjEQ_", "mjtWrap": "mjWRAP_", "mjtIntegrator": "mjINT_",
  "mjtCone": "mjCONE_", "mjtSolver": "mjSOL_", "mjtJacobian": "mjJAC_", "mjtConstraint": "mjCNSTR_", "mjtConstraintState": "mjCNSTRSTATE_",
  "mjtCamLight": "mjCAMLIGHT_", "mjtStage": "mjSTAGE_", "mjtDataType": "mjDATATYPE_", "mjtProjection": "mjPROJ_", "mjtSleepPolicy": "mjSLEEP_",
  "mjtSleepState": "mjS_", "mjtTextureRole": "mjTEXROLE_", "mjtTexture": "mjTEXTURE_", "mjtLightType": "mjLIGHT_", "mjtCamOutBit": "mjCAMOUT_",
}
# not on the step path: members only need distinct values
_SENSORS = """TOUCH ACCELEROMETER VELOCIMETER GYRO FORCE TORQUE MAGNETOMETER RANGEFINDER CAMPROJECTION JOINTPOS JOINTVEL TENDONPOS TENDONVEL
ACTUATORPOS ACTUATORVEL ACTUATORFRC JOINTACTFRC TENDONACTFRC BALLQUAT BALLANGVEL JOINTLIMITPOS JOINTLIMITVEL JOINTLIMITFRC TENDONLIMITPOS
TENDONLIMITVEL TENDONLIMITFRC FRAMEPOS FRAMEQUAT FRAMEXAXIS FRAMEYAXIS FRAMEZAXIS FRAMELINVEL FRAMEANGVEL FRAMELINACC FRAMEANGACC SUBTREECOM
SUBTREELINVEL SUBTREEANGMOM INSIDESITE GEOMDIST GEOMNORMAL GEOMFROMTO CONTACT E_POTENTIAL E_KINETIC CLOCK TACTILE PLUGIN USER""".split()
_STATES = "TIME QPOS QVEL ACT HISTORY WARMSTART CTRL QFRC_APPLIED XFRC_APPLIED EQ_ACTIVE MOCAP_POS MOCAP_QUAT USERDATA PLUGIN".split()


def fake_mujoco():
  mj = _t.ModuleType("mujoco")
  for ename, members in _ENUMS.items():
    cls = enum.IntEnum(ename, {_PREFIX[ename] + k: v for k, v in members.items()})
    setattr(mj, ename, cls)
  mj.mjtTextureRole.mjNTEXROLE = 10
  mj.mjtSensor = enum.IntEnum("mjtSensor", {"mjSENS_" + k: i for i, k in enumerate(_SENSORS)})
  st = {"mjSTATE_" + k: 1 << i for i, k in enumerate(_STATES)}
  st["mjSTATE_PHYSICS"] = st["mjSTATE_QPOS"] | st["mjSTATE_QVEL"] | st["mjSTATE_ACT"] | st["mjSTATE_HISTORY"]
  st["mjSTATE_FULLPHYSICS"] = st["mjSTATE_TIME"] | st["mjSTATE_PHYSICS"] | st["mjSTATE_PLUGIN"]
  st["mjSTATE_USER"] = st["mjSTATE_CTRL"] | st["mjSTATE_QFRC_APPLIED"] | st["mjSTATE_XFRC_APPLIED"] | st["mjSTATE_EQ_ACTIVE"] | st["mjSTATE_MOCAP_POS"] | st["mjSTATE_MOCAP_QUAT"] | st["mjSTATE_USERDATA"]
  st["mjSTATE_INTEGRATION"] = st["mjSTATE_FULLPHYSICS"] | st["mjSTATE_USER"] | st["mjSTATE_WARMSTART"]
  mj.mjtState = enum.IntEnum("mjtState", st)
  mj.mjtState.mjNSTATE = len(_STATES)
  mj.mjMINVAL, mj.mjMAXVAL, mj.mjMINIMP, mj.mjMAXIMP, mj.mjMINMU, mj.mjMAXCONPAIR, mj.mjMINAWAKE = 1e-15, 1e10, 0.0001, 0.9999, 1e-5, 50, 10
  mj.mjNPOLY, mj.mjNFLUID, mj.mjNEQDATA, mj.mjNGAIN, mj.mjNBIAS, mj.mjNDYN, mj.mjNIMP, mj.mjNREF = 2, 12, 11, 10, 10, 10, 5, 2
  mj.MjModel = type("MjModel", (), {})
  class MjData:
    """xquat / xmat / ximat / geom_xpos / geom_xmat at qpos0 (what io.make_data reads, io.py:1818-1840), computed by the
    mjcf compiler's host kinematics; mj_kinematics is then a no-op"""

    def __init__(self, m):
      from mujoco_warp_b200._src import mjcf

      base = object.__getattribute__(m, "_m") if isinstance(m, MjModelAdapter) else m
      kin = mjcf.kinematics_np(base, base.qpos0)
      for k in ("xpos", "xquat", "xmat", "xipos", "ximat", "geom_xpos", "geom_xmat"):
        if hasattr(kin, k):
          a = np.asarray(getattr(kin, k), dtype=np.float64)
          setattr(self, k, a.reshape(a.shape[0], -1))  # MjData stores matrices flat: (n, 9)
      self.nisland, self.nidof = 0, 0
      self.tree_island = -np.ones(base.ntree, dtype=np.int32)
      self.dof_island = -np.ones(base.nv, dtype=np.int32)

  mj.MjData = MjData
  mj.mj_kinematics = lambda m, d: None
  mj.MjSpec = type("MjSpec", (), {})
  mj.mj_isSparse = lambda m: int(m.opt.jacobian == 1 or (m.opt.jacobian == 2 and m.nv >= 60))
  mj.mj_name2id = lambda m, t, n: -1
  mj.mj_id2name = lambda m, t, i: None
  mj.__version__ = "fake"
  sys.modules["mujoco"] = mj
  return mj


class MjModelAdapter:
  """MjModel-shaped view of a model compiled by mujoco_warp_b200._src.mjcf.  Fields the compiler does not produce (sensors,
  tendons, flexes, plugins, ...) read as empty arrays / zero counts; every such fallback is recorded in `.missing`."""

  def __init__(self, mjm, defaults=None):
    object.__setattr__(self, "_m", mjm)
    object.__setattr__(self, "missing", set())
    d = dict(defaults or {})
    # MjModel sizes the compiler does not store: non-zeros of the actuator moment rows (joint transmissions)
    nnz = {0: 6, 1: 3, 2: 1, 3: 1}
    trn = np.asarray(mjm.actuator_trnid).reshape(-1, 2)
    d.setdefault("nJmom", int(sum(nnz[int(mjm.jnt_type[j])] for j in trn[:, 0])) if mjm.nu else 0)
    # the mjcf compiler stores each exclude in both body orders; MjModel holds one (sorted) signature per exclude
    if len(getattr(mjm, "exclude_signature", ())):
      d.setdefault("exclude_signature", np.asarray(mjm.exclude_signature)[::2])
    object.__setattr__(self, "_defaults", d)

  def __getattr__(self, name):
    m = object.__getattribute__(self, "_m")
    d = object.__getattribute__(self, "_defaults")
    if name in d:
      return d[name]
    if name == "opt":
      return OptAdapter(m.opt)
    if hasattr(m, name):
      v = getattr(m, name)
      return v
    self.missing.add(name)
    if name.startswith("n") and "_" not in name:
      return 0
    v = self._from_spec(name)
    object.__getattribute__(self, "_defaults")[name] = v
    return v

  def _from_spec(self, name):
    """zeros shaped like the reference's Model field spec (types.py `array("ngeom", 3, float)` ...); *_plugin ids are -1"""
    import dataclasses

    ty = sys.modules["mujoco_warp._src.types"]
    names = [f.name for f in dataclasses.fields(ty.Model)]
    if name in names and names.index(name) >= names.index("callback"):
      raise AttributeError(name)  # derived ("warp only", types.py:860) fields are computed by put_model, not read from MjModel
    spec = {f.name: f.type for f in dataclasses.fields(ty.Model)}.get(name)
    shape = getattr(spec, "shape", None)
    if not shape:
      return np.zeros((0,), dtype=np.int32)
    dims = []
    for d in shape:
      if d == "*":
        continue
      if isinstance(d, str):
        d = int(d) if d.isdigit() else int(getattr(self, d))
      dims.append(int(d))
    inner = tuple(getattr(spec.dtype, "_shape_", ()))
    npdt = warp_shim._np_dtype(spec.dtype)
    out = np.zeros(tuple(dims) + inner, dtype=npdt)
    if name.endswith("_plugin"):
      out[...] = -1
    return out


OPT_DEFAULTS = dict(  # MuJoCo mjOption defaults for fields the mjcf compiler does not carry
  noslip_iterations=0, noslip_tolerance=1e-6, ccd_iterations=35, ccd_tolerance=1e-6, sdf_iterations=10, sdf_initpoints=40,
  jacobian=0,  # mjJAC_DENSE: this framework keeps efc.J dense for every supported model (nv <= 64); AUTO would switch to sparse above nv = 32
  o_margin=0.0, o_solref=np.array([0.02, 1.0]), o_solimp=np.array([0.9, 0.95, 0.001, 0.5, 2.0]), o_friction=np.array([1, 1, 0.005, 0.0001, 0.0001]),
  wind=np.zeros(3), magnetic=np.array([0, -0.5, 0.0]), density=0.0, viscosity=0.0, sleep_tolerance=1e-4, disableactuator=0, apirate=100.0,
)


class OptAdapter:
  def __init__(self, opt):
    self._o = opt

  def __getattr__(self, name):
    o = object.__getattribute__(self, "_o")
    if name == "jacobian":  # dense unless the model asks otherwise (the reference refuses dense above nv = 60)
      return int(getattr(o, "jacobian", OPT_DEFAULTS["jacobian"]))
    if hasattr(o, name):
      return getattr(o, name)
    if name in OPT_DEFAULTS:
      return OPT_DEFAULTS[name]
    raise AttributeError(f"opt.{name}")


def setup():
  """install the shim + fake mujoco and import the reference modules needed for the step path"""
  fake_mujoco()  # before install(): the real types.py is then imported against these enum values
  wp = warp_shim.install()
  ref = {}
  for name in ("types", "io", "forward", "smooth", "collision_driver", "constraint", "solver", "passive", "support", "math"):
    ref[name] = warp_shim.load_reference_module(name)
  return wp, ref


if __name__ == "__main__":
  wp, ref = setup()
  print("loaded", sorted(ref))
