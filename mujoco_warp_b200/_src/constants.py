"""MuJoCo constants and enum values used on the step path.

The reference reads these from the `mujoco` module at import time
(`/root/reference/mujoco_warp/_src/types.py:23-29,260-277,413-544`).  That package is not
available here, so the values below are restated from the public MuJoCo headers (mjmodel.h).
They are assumptions in the sense of SURVEY.md §7 step 0: self-consistency is unit-tested
(tests/test_constants.py); when `mujoco` is importable the test also cross-checks them.
"""

MJ_MINVAL = 1e-15
MJ_MAXVAL = 1e10
MJ_MINIMP = 1e-4
MJ_MAXIMP = 0.9999
MJ_MAXCONPAIR = 50
MJ_MINMU = 1e-5

# mjtJoint
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
# mjtGeom
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH, GEOM_SDF = range(9)
NGEOMTYPES = 10  # len(GeomType) incl. FLEX in the reference enum (types.py:453-480), used for pair-type tables
# mjtIntegrator
INT_EULER, INT_RK4, INT_IMPLICIT, INT_IMPLICITFAST = 0, 1, 2, 3
# mjtCone
CONE_PYRAMIDAL, CONE_ELLIPTIC = 0, 1
# mjtJacobian
JAC_DENSE, JAC_SPARSE, JAC_AUTO = 0, 1, 2
# mjtSolver
SOL_PGS, SOL_CG, SOL_NEWTON = 0, 1, 2
# mjtConstraint
CNSTR_EQUALITY, CNSTR_FRICTION_DOF, CNSTR_FRICTION_TENDON, CNSTR_LIMIT_JOINT, CNSTR_LIMIT_TENDON = 0, 1, 2, 3, 4
CNSTR_CONTACT_FRICTIONLESS, CNSTR_CONTACT_PYRAMIDAL, CNSTR_CONTACT_ELLIPTIC = 5, 6, 7
# mjtConstraintState
STATE_SATISFIED, STATE_QUADRATIC, STATE_LINEARNEG, STATE_LINEARPOS, STATE_CONE = 0, 1, 2, 3, 4
# mjtTrn / mjtDyn / mjtGain / mjtBias
TRN_JOINT, TRN_JOINTINPARENT, TRN_SLIDERCRANK, TRN_TENDON, TRN_SITE, TRN_BODY = range(6)
DYN_NONE = 0
GAIN_FIXED, GAIN_AFFINE, GAIN_MUSCLE = 0, 1, 2
BIAS_NONE, BIAS_AFFINE, BIAS_MUSCLE = 0, 1, 2
# mjtCamLight
CAMLIGHT_FIXED, CAMLIGHT_TRACK, CAMLIGHT_TRACKCOM, CAMLIGHT_TARGETBODY, CAMLIGHT_TARGETBODYCOM = range(5)
CAMLIGHT_MODES = {"fixed": 0, "track": 1, "trackcom": 2, "targetbody": 3, "targetbodycom": 4}

# mjtDisableBit
DSBL_CONSTRAINT = 1 << 0
DSBL_EQUALITY = 1 << 1
DSBL_FRICTIONLOSS = 1 << 2
DSBL_LIMIT = 1 << 3
DSBL_CONTACT = 1 << 4
DSBL_SPRING = 1 << 5
DSBL_DAMPER = 1 << 6
DSBL_GRAVITY = 1 << 7
DSBL_CLAMPCTRL = 1 << 8
DSBL_WARMSTART = 1 << 9
DSBL_FILTERPARENT = 1 << 10
DSBL_ACTUATION = 1 << 11
DSBL_REFSAFE = 1 << 12
DSBL_SENSOR = 1 << 13
DSBL_MIDPHASE = 1 << 14
DSBL_EULERDAMP = 1 << 15
DSBL_AUTORESET = 1 << 16
DSBL_NATIVECCD = 1 << 17
DSBL_ISLAND = 1 << 18
DISABLE_FLAGS = {
  "constraint": DSBL_CONSTRAINT,
  "equality": DSBL_EQUALITY,
  "frictionloss": DSBL_FRICTIONLOSS,
  "limit": DSBL_LIMIT,
  "contact": DSBL_CONTACT,
  "spring": DSBL_SPRING,
  "damper": DSBL_DAMPER,
  "gravity": DSBL_GRAVITY,
  "clampctrl": DSBL_CLAMPCTRL,
  "warmstart": DSBL_WARMSTART,
  "filterparent": DSBL_FILTERPARENT,
  "actuation": DSBL_ACTUATION,
  "refsafe": DSBL_REFSAFE,
  "sensor": DSBL_SENSOR,
  "midphase": DSBL_MIDPHASE,
  "eulerdamp": DSBL_EULERDAMP,
  "autoreset": DSBL_AUTORESET,
  "nativeccd": DSBL_NATIVECCD,
  "island": DSBL_ISLAND,
}
ENBL_OVERRIDE, ENBL_ENERGY, ENBL_FWDINV, ENBL_INVDISCRETE, ENBL_MULTICCD, ENBL_SLEEP = (1 << i for i in range(6))
ENABLE_FLAGS = {"override": 1, "energy": 2, "fwdinv": 4, "invdiscrete": 8, "multiccd": 16, "sleep": 32}

# OverflowType (reference types.py:149-176; warp-only, so exact)
OVF_NEFC = 1 << 0
OVF_NJMAX_NNZ = 1 << 1
OVF_BROADPHASE = 1 << 2
OVF_NARROWPHASE = 1 << 3
OVF_CCD = 1 << 4
OVF_ITERATIONS = 1 << 9
OVF_LS_ITERATIONS = 1 << 10

# BroadphaseFilter (reference types.py:101-118)
BF_PLANE, BF_SPHERE, BF_AABB, BF_OBB = 1, 2, 4, 8
# BroadphaseType
BROADPHASE_NXN, BROADPHASE_SAP_TILE, BROADPHASE_SAP_SEGMENTED = 0, 1, 2
# ContactType (reference types.py)
CONTACT_TYPE_CONSTRAINT, CONTACT_TYPE_SENSOR = 1, 2

DEFAULT_SOLREF = [0.02, 1.0]
DEFAULT_SOLIMP = [0.9, 0.95, 0.001, 0.5, 2.0]
DEFAULT_FRICTION = [1.0, 0.005, 0.0001]

# mjtEq / mjtObj (the members the equality path uses)
EQ_CONNECT, EQ_WELD, EQ_JOINT, EQ_TENDON, EQ_FLEX = 0, 1, 2, 3, 4
OBJ_BODY, OBJ_JOINT, OBJ_SITE = 1, 3, 6
