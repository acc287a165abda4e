/* oracle.c -- TEST INFRASTRUCTURE ONLY. CPU restatement of mujoco_warp's batched step.
 * See oracle.h for scope and the parity-pinning statement. Reference paths are relative to
 * /root/reference/mujoco_warp/_src/.  Layout conventions: quaternions (w,x,y,z),
 * spatial vectors (angular[3], linear[3]), vec10 inertia = [Ixx,Iyy,Izz,Ixy,Ixz,Iyz, m*dx,m*dy,m*dz, m],
 * mat33 row-major, contact frame row 0 = normal (geom1 -> geom2).
 */
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MJ_MINVAL ((real)1e-15)
#define MJ_MAXVAL ((real)1e10)
#define MJ_MINIMP ((real)1e-4)
#define MJ_MAXIMP ((real)0.9999)
#define MJ_MINMU ((real)1e-5)

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { GEOM_PLANE = 0, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH };
enum { INT_EULER = 0, INT_RK4, INT_IMPLICIT, INT_IMPLICITFAST };
enum { CONE_PYRAMIDAL = 0, CONE_ELLIPTIC = 1 };
enum { SOL_CG = 1, SOL_NEWTON = 2 };
enum { EQ_CONNECT = 0, EQ_WELD = 1, EQ_JOINT = 2, EQ_TENDON = 3 };
enum { TRN_JOINT = 0, TRN_TENDON = 3 };
enum { CNSTR_EQUALITY = 0, CNSTR_FRICTION_DOF = 1, CNSTR_FRICTION_TENDON = 2, CNSTR_LIMIT_JOINT = 3, CNSTR_LIMIT_TENDON = 4, CNSTR_CONTACT_FRICTIONLESS = 5, CNSTR_CONTACT_PYRAMIDAL = 6, CNSTR_CONTACT_ELLIPTIC = 7 };
enum { ST_SATISFIED = 0, ST_QUADRATIC = 1, ST_LINEARNEG = 2, ST_LINEARPOS = 3, ST_CONE = 4 };
enum { CAM_FIXED = 0, CAM_TRACK, CAM_TRACKCOM, CAM_TARGETBODY, CAM_TARGETBODYCOM };
enum { GAIN_FIXED = 0, GAIN_AFFINE = 1 };
enum { BIAS_NONE = 0, BIAS_AFFINE = 1 };
enum {
  DSBL_CONSTRAINT = 1 << 0, DSBL_EQUALITY = 1 << 1, DSBL_FRICTIONLOSS = 1 << 2, DSBL_LIMIT = 1 << 3, DSBL_CONTACT = 1 << 4,
  DSBL_SPRING = 1 << 5, DSBL_DAMPER = 1 << 6, DSBL_GRAVITY = 1 << 7, DSBL_CLAMPCTRL = 1 << 8, DSBL_WARMSTART = 1 << 9,
  DSBL_ACTUATION = 1 << 11, DSBL_REFSAFE = 1 << 12, DSBL_EULERDAMP = 1 << 15, DSBL_NATIVECCD = 1 << 17
};
enum { OVF_NEFC = 1 << 0, OVF_NARROWPHASE = 1 << 3, OVF_CCD = 1 << 4, OVF_EPA_HORIZON = 1 << 8, OVF_ITERATIONS = 1 << 9, OVF_LS_ITERATIONS = 1 << 10, OVF_UNSUPPORTED = 1 << 30 };
enum { BF_PLANE = 1, BF_SPHERE = 2, BF_AABB = 4, BF_OBB = 8 };

/* ------------------------------------------------------------------ field registries (X-macros) */
#define MODEL_INTS(X) \
  X(nq) X(nv) X(nu) X(nbody) X(nmocap) X(njnt) X(ngeom) X(nsite) X(ncam) X(nlight) X(nC) X(ntree) X(nJmom) \
  X(nxn_npair) X(nlimit) X(nlimit_ball) X(neq) X(nmaxpyramid) X(integrator) X(cone) X(solver) X(iterations) X(ls_iterations) \
  X(disableflags) X(enableflags) X(broadphase) X(broadphase_filter) X(ccd_iterations) X(epa_iterations) X(nsensor) X(nsensordata) X(nmesh) X(na) X(ntendon) X(nJten)
#define MODEL_REALS(X) X(timestep) X(tolerance) X(ls_tolerance) X(impratio_invsqrt) X(meaninertia) X(ccd_tolerance)
#define MODEL_IARRS(X) \
  X(body_parentid) X(body_rootid) X(body_weldid) X(body_mocapid) X(body_jntnum) X(body_jntadr) X(body_dofnum) X(body_dofadr) \
  X(jnt_type) X(jnt_qposadr) X(jnt_dofadr) X(jnt_bodyid) X(jnt_actfrclimited) X(jnt_actgravcomp) \
  X(dof_bodyid) X(dof_jntid) X(dof_parentid) X(M_rownnz) X(M_rowadr) X(M_colind) \
  X(tree_dofadr) X(tree_dofnum) X(qLD_block_adr) \
  X(geom_type) X(geom_condim) X(geom_bodyid) X(geom_priority) \
  X(actuator_trnid) X(actuator_gaintype) X(actuator_biastype) X(actuator_ctrllimited) X(actuator_forcelimited) \
  X(actuator_dyntype) X(actuator_actadr) X(actuator_actnum) X(actuator_actlimited) X(actuator_actearly) X(actuator_trntype) \
  X(ten_J_rownnz) X(ten_J_rowadr) X(ten_J_colind) X(tendon_adr) X(tendon_num) X(wrap_objid) X(tendon_limited) X(tendon_actfrclimited) \
  X(cam_mode) X(cam_bodyid) X(cam_targetbodyid) X(light_mode) X(light_bodyid) X(light_targetbodyid) X(site_bodyid) \
  X(nxn_geom_pair) X(nxn_pairid) X(jnt_limited_slide_hinge_adr) X(jnt_limited_ball_adr) X(body_isdofancestor) \
  X(eq_type) X(eq_obj1id) X(eq_obj2id) X(pair_dim) \
  X(sensor_type) X(sensor_datatype) X(sensor_needstage) X(sensor_objtype) X(sensor_objid) X(sensor_reftype) X(sensor_refid) X(sensor_dim) X(sensor_adr) X(site_type) \
  X(geom_dataid) X(mesh_vertadr) X(mesh_vertnum) X(mesh_graphadr) X(mesh_graph) X(mesh_polynum) X(mesh_polyadr) X(mesh_polyvertadr) X(mesh_polyvertnum) \
  X(mesh_polyvert) X(mesh_polymapadr) X(mesh_polymapnum) X(mesh_polymap)
#define MODEL_RARRS(X) \
  X(gravity) X(qpos0) X(qpos_spring) X(body_pos) X(body_quat) X(body_ipos) X(body_iquat) X(body_mass) X(body_subtreemass) \
  X(body_inertia) X(body_invweight0) X(body_gravcomp) X(jnt_pos) X(jnt_axis) X(jnt_stiffness) X(jnt_range) X(jnt_margin) X(jnt_solref) \
  X(jnt_solimp) X(jnt_actfrcrange) X(dof_armature) X(dof_damping) X(dof_invweight0) X(dof_frictionloss) X(dof_solref) \
  X(dof_solimp) X(geom_size) X(geom_aabb) X(geom_rbound) X(geom_pos) X(geom_quat) X(geom_friction) X(geom_margin) \
  X(geom_gap) X(geom_solmix) X(geom_solref) X(geom_solimp) X(actuator_gear) X(actuator_gainprm) X(actuator_biasprm) \
  X(actuator_ctrlrange) X(actuator_forcerange) X(cam_pos) X(cam_quat) X(cam_poscom0) X(cam_pos0) X(cam_mat0) \
  X(light_pos) X(light_dir) X(light_poscom0) X(light_pos0) X(light_dir0) X(site_pos) X(site_quat) \
  X(eq_solref) X(eq_solimp) X(eq_data) X(pair_friction) X(pair_solref) X(pair_solreffriction) X(pair_solimp) X(pair_margin) X(pair_gap) \
  X(sensor_cutoff) X(site_size) X(mesh_vert) X(mesh_polynormal) X(actuator_dynprm) X(actuator_actrange) \
  X(wrap_prm) X(tendon_range) X(tendon_margin) X(tendon_stiffness) X(tendon_damping) X(tendon_frictionloss) X(tendon_lengthspring) \
  X(tendon_length0) X(tendon_invweight0) X(tendon_solref_lim) X(tendon_solimp_lim) X(tendon_solref_fri) X(tendon_solimp_fri) X(tendon_actfrcrange)

/* Data arrays: (nworld, per-world size) row-major; per-world sizes are implied by the model dims. */
#define DATA_RARRS(X) \
  X(time) X(qpos) X(qvel) X(ctrl) X(qacc_warmstart) X(qfrc_applied) X(xfrc_applied) X(qacc) X(mocap_pos) X(mocap_quat) \
  X(xpos) X(xquat) X(xmat) X(xipos) X(ximat) X(xanchor) X(xaxis) X(geom_xpos) X(geom_xmat) X(site_xpos) X(site_xmat) \
  X(cam_xpos) X(cam_xmat) X(light_xpos) X(light_xdir) X(subtree_com) X(cdof) X(cinert) X(crb) X(M) X(qLD) \
  X(actuator_length) X(actuator_moment) X(actuator_velocity) X(cvel) X(cdof_dot) X(qfrc_bias) X(qfrc_spring) \
  X(qfrc_damper) X(qfrc_gravcomp) X(qfrc_passive) X(actuator_force) X(qfrc_actuator) X(qfrc_smooth) X(qacc_smooth) \
  X(qfrc_constraint) X(cacc) X(cfrc_int) X(cfrc_ext) X(sensordata) X(subtree_linvel) X(subtree_angmom) \
  X(efc_J) X(efc_pos) X(efc_margin) X(efc_D) X(efc_vel) X(efc_aref) X(efc_frictionloss) X(efc_force) X(efc_Ma) \
  X(con_dist) X(con_pos) X(con_frame) X(con_includemargin) X(con_friction) X(con_solref) X(con_solreffriction) X(con_solimp) \
  X(act) X(act_dot) X(ten_length) X(ten_J) X(ten_velocity)
#define DATA_IARRS(X) \
  X(ne) X(nf) X(nl) X(nefc) X(ncon) X(ncollision) X(solver_niter) X(overflow) X(efc_type) X(efc_id) X(efc_state) \
  X(moment_rownnz) X(moment_rowadr) X(moment_colind) X(con_dim) X(con_geom) X(con_efc_address) X(con_geomcollisionid) \
  X(eq_active)

struct OrcModel {
#define X(n) int n;
  MODEL_INTS(X)
#undef X
#define X(n) real n;
  MODEL_REALS(X)
#undef X
#define X(n) const int* n;
  MODEL_IARRS(X)
#undef X
#define X(n) const real* n;
  MODEL_RARRS(X)
#undef X
};

struct OrcData {
  int nworld, nconmax, njmax;
#define X(n) real* n;
  DATA_RARRS(X)
#undef X
#define X(n) int* n;
  DATA_IARRS(X)
#undef X
};

static char g_err[256] = "";
const char* orc_last_error(void) { return g_err; }
int orc_sizeof_real(void) { return (int)sizeof(real); }

OrcModel* orc_model_create(void) { return (OrcModel*)calloc(1, sizeof(OrcModel)); }
void orc_model_free(OrcModel* m) { free(m); }
int orc_model_set_int(OrcModel* m, const char* name, int v) {
#define X(n) if (!strcmp(name, #n)) { m->n = v; return 0; }
  MODEL_INTS(X)
#undef X
  return -1;
}
int orc_model_set_real(OrcModel* m, const char* name, double v) {
#define X(n) if (!strcmp(name, #n)) { m->n = (real)v; return 0; }
  MODEL_REALS(X)
#undef X
  return -1;
}
int orc_model_set_iarr(OrcModel* m, const char* name, const int* p) {
#define X(n) if (!strcmp(name, #n)) { m->n = p; return 0; }
  MODEL_IARRS(X)
#undef X
  return -1;
}
int orc_model_set_rarr(OrcModel* m, const char* name, const real* p) {
#define X(n) if (!strcmp(name, #n)) { m->n = p; return 0; }
  MODEL_RARRS(X)
#undef X
  return -1;
}
OrcData* orc_data_create(int nworld, int nconmax, int njmax) {
  OrcData* d = (OrcData*)calloc(1, sizeof(OrcData));
  d->nworld = nworld; d->nconmax = nconmax; d->njmax = njmax;
  return d;
}
void orc_data_free(OrcData* d) { free(d); }
int orc_data_set_iarr(OrcData* d, const char* name, int* p) {
#define X(n) if (!strcmp(name, #n)) { d->n = p; return 0; }
  DATA_IARRS(X)
#undef X
  return -1;
}
int orc_data_set_rarr(OrcData* d, const char* name, real* p) {
#define X(n) if (!strcmp(name, #n)) { d->n = p; return 0; }
  DATA_RARRS(X)
#undef X
  return -1;
}

static int check_fields(const OrcModel* m, const OrcData* d) {
#define X(n) if (!m->n) { snprintf(g_err, sizeof g_err, "model array '%s' not set", #n); return -1; }
  MODEL_IARRS(X)
  MODEL_RARRS(X)
#undef X
#define X(n) if (!d->n) { snprintf(g_err, sizeof g_err, "data array '%s' not set", #n); return -1; }
  DATA_RARRS(X)
  DATA_IARRS(X)
#undef X
  return 0;
}

/* ------------------------------------------------------------------ math (math.py) */
static inline real rmin(real a, real b) { return a < b ? a : b; }
static inline real rmax(real a, real b) { return a > b ? a : b; }
static inline real rclamp(real x, real lo, real hi) { return rmin(rmax(x, lo), hi); }
static inline real dot3(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(const real* a, const real* b, real* o) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline real len3(const real* a) { return (real)sqrt((double)dot3(a, a)); }
/* wp.normalize: zero vector stays zero */
static inline void normalize3(real* a) { real l = len3(a); if (l > 0) { a[0] /= l; a[1] /= l; a[2] /= l; } else { a[0] = a[1] = a[2] = 0; } }
static inline void normalize4(real* q) {
  real l = (real)sqrt((double)(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]));
  if (l > 0) { q[0] /= l; q[1] /= l; q[2] /= l; q[3] /= l; } else { q[0] = q[1] = q[2] = q[3] = 0; }
}
/* math.py:24 */
static inline void mul_quat(const real* u, const real* v, real* o) {
  real r0 = u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3];
  real r1 = u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2];
  real r2 = u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1];
  real r3 = u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0];
  o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3;
}
/* math.py:45 */
static inline void rot_vec_quat(const real* v, const real* q, real* o) {
  real s = q[0]; const real* u = q + 1;
  real uv = dot3(u, v), uu = dot3(u, u), c[3];
  cross3(u, v, c);
  for (int i = 0; i < 3; i++) o[i] = 2 * uv * u[i] + (s * s - uu) * v[i] + 2 * s * c[i];
}
/* math.py:53 */
static inline void axis_angle_to_quat(const real* axis, real angle, real* q) {
  real s = (real)sin((double)angle * 0.5), c = (real)cos((double)angle * 0.5);
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
/* math.py:60 */
static inline void quat_to_mat(const real* q, real* m) {
  real q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  real q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3], q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[1] = 2 * (q12 - q03); m[2] = 2 * (q13 + q02);
  m[3] = 2 * (q12 + q03); m[4] = q00 - q11 + q22 - q33; m[5] = 2 * (q23 - q01);
  m[6] = 2 * (q13 - q02); m[7] = 2 * (q23 + q01); m[8] = q00 - q11 - q22 + q33;
}
/* math.py:121 mju_mulInertVec */
static inline void inert_vec(const real* i, const real* v, real* o) {
  real r[6];
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
  memcpy(o, r, sizeof r);
}
/* math.py:134 */
static inline void motion_cross(const real* u, const real* v, real* o) {
  real a[3], b[3], c[3];
  cross3(u, v, a); cross3(u + 3, v, b); cross3(u, v + 3, c);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
/* math.py:148 */
static inline void motion_cross_force(const real* v, const real* f, real* o) {
  real a[3], b[3], c[3];
  cross3(v, f, a); cross3(v + 3, f + 3, b); cross3(v, f + 3, c);
  o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}
/* math.py:189 */
static inline void quat_integrate(real* q, const real* v, real dt) {
  real ax[3] = {v[0], v[1], v[2]};
  real n = len3(ax);
  normalize3(ax);
  real qr[4], out[4];
  axis_angle_to_quat(ax, dt * n, qr);
  normalize4(q);
  mul_quat(q, qr, out);
  normalize4(out);
  memcpy(q, out, sizeof out);
}
/* math.py:203 orthogonals + math.py:247 make_frame (rows: a, b, c) */
static inline void make_frame(const real* a_in, real* frame) {
  real a[3] = {a_in[0], a_in[1], a_in[2]};
  normalize3(a);
  real y[3] = {0, 1, 0}, z[3] = {0, 0, 1}, b[3];
  const real* s = (-0.5 < a[1] && a[1] < 0.5) ? y : z;
  real ab = dot3(a, s);
  for (int i = 0; i < 3; i++) b[i] = s[i] - a[i] * ab;
  normalize3(b);
  if (len3(a) == 0) b[0] = b[1] = b[2] = 0;
  real c[3];
  cross3(a, b, c);
  for (int i = 0; i < 3; i++) { frame[i] = a[i]; frame[3 + i] = b[i]; frame[6 + i] = c[i]; }
}
static inline real safe_div(real x, real y) { return x / (y != 0 ? y : MJ_MINVAL); }
/* math.py:269 */
void orc_closest_segment_point(const real a[3], const real b[3], const real pt[3], real out[3]) {
  real ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, pa[3] = {pt[0] - a[0], pt[1] - a[1], pt[2] - a[2]};
  real t = dot3(pa, ab) / (dot3(ab, ab) + (real)1e-6);
  t = rclamp(t, 0, 1);
  for (int i = 0; i < 3; i++) out[i] = a[i] + t * ab[i];
}
static real seg_pt_dist(const real a[3], const real b[3], const real pt[3], real out[3]) {
  orc_closest_segment_point(a, b, pt, out);
  real d[3] = {pt[0] - out[0], pt[1] - out[1], pt[2] - out[2]};
  return dot3(d, d);
}
/* math.py:285 */
void orc_closest_segment_to_segment_points(const real a0[3], const real a1[3], const real b0[3], const real b1[3], real outa[3], real outb[3]) {
  real da[3], db[3];
  for (int i = 0; i < 3; i++) { da[i] = a1[i] - a0[i]; db[i] = b1[i] - b0[i]; }
  real la = len3(da), lb = len3(db);
  if (la != 0) for (int i = 0; i < 3; i++) da[i] /= la;
  if (lb != 0) for (int i = 0; i < 3; i++) db[i] /= lb;
  real ha = la * (real)0.5, hb = lb * (real)0.5, am[3], bm[3], tr[3];
  for (int i = 0; i < 3; i++) { am[i] = a0[i] + da[i] * ha; bm[i] = b0[i] + db[i] * hb; tr[i] = am[i] - bm[i]; }
  real dd = dot3(da, db), dat = dot3(da, tr), dbt = dot3(db, tr), denom = 1 - dd * dd;
  real ta = (-dat + dd * dbt) / (denom + (real)1e-6), tb = dbt + ta * dd;
  ta = rclamp(ta, -ha, ha); tb = rclamp(tb, -hb, hb);
  real ba[3], bb[3], na[3], nb[3];
  for (int i = 0; i < 3; i++) { ba[i] = am[i] + da[i] * ta; bb[i] = bm[i] + db[i] * tb; }
  real d1 = seg_pt_dist(a0, a1, bb, na), d2 = seg_pt_dist(b0, b1, ba, nb);
  if (d1 < d2) { memcpy(outa, na, 3 * sizeof(real)); memcpy(outb, bb, 3 * sizeof(real)); }
  else { memcpy(outa, ba, 3 * sizeof(real)); memcpy(outb, nb, 3 * sizeof(real)); }
}
int orc_upper_tri_index(int n, int i, int j) { return (i * (2 * n - i - 3)) / 2 + j - 1; }
int orc_upper_trid_index(int n, int i, int j) { if (j < i) { int t = i; i = j; j = t; } return (i * (2 * n - i - 1)) / 2 + j; }
/* util_misc.py:61 */
double orc_halton(int index, int base) {
  int n0 = index; float b = (float)base, f = 1.0f / b, hn = 0.0f;
  while (n0 > 0) { int n1 = n0 / base; int r = n0 - n1 * base; hn += f * (float)r; f /= b; n0 = n1; }
  return (double)hn;
}

/* ------------------------------------------------------------------ per-world view */
typedef struct {
  const OrcModel* m;
  int njmax, nconmax;
#define X(n) real* n;
  DATA_RARRS(X)
#undef X
#define X(n) int* n;
  DATA_IARRS(X)
#undef X
} W;

static void make_view(const OrcModel* m, const OrcData* d, int w, W* v) {
  const int nb = m->nbody, nv = m->nv, nj = m->njnt, ng = m->ngeom, nu = m->nu, njm = d->njmax, ncm = d->nconmax;
  v->m = m; v->njmax = njm; v->nconmax = ncm;
#define R(n, sz) v->n = d->n + (size_t)w * (size_t)(sz)
  R(mocap_pos, 3 * m->nmocap); R(mocap_quat, 4 * m->nmocap);
  R(time, 1); R(qpos, m->nq); R(qvel, nv); R(ctrl, nu); R(qacc_warmstart, nv); R(qfrc_applied, nv); R(xfrc_applied, 6 * nb); R(qacc, nv);
  R(xpos, 3 * nb); R(xquat, 4 * nb); R(xmat, 9 * nb); R(xipos, 3 * nb); R(ximat, 9 * nb); R(xanchor, 3 * nj); R(xaxis, 3 * nj);
  R(geom_xpos, 3 * ng); R(geom_xmat, 9 * ng); R(site_xpos, 3 * m->nsite); R(site_xmat, 9 * m->nsite);
  R(cam_xpos, 3 * m->ncam); R(cam_xmat, 9 * m->ncam); R(light_xpos, 3 * m->nlight); R(light_xdir, 3 * m->nlight);
  R(subtree_com, 3 * nb); R(cdof, 6 * nv); R(cinert, 10 * nb); R(crb, 10 * nb); R(M, m->nC);
  {
    int qld = 0;
    for (int t = 0; t < m->ntree; t++) qld += m->tree_dofnum[t] * m->tree_dofnum[t];
    R(qLD, qld);
  }
  R(actuator_length, nu); R(actuator_moment, m->nJmom); R(actuator_velocity, nu); R(cvel, 6 * nb); R(cdof_dot, 6 * nv);
  R(qfrc_bias, nv); R(qfrc_spring, nv); R(qfrc_damper, nv); R(qfrc_gravcomp, nv); R(qfrc_passive, nv); R(actuator_force, nu);
  R(qfrc_actuator, nv); R(qfrc_smooth, nv); R(qacc_smooth, nv); R(qfrc_constraint, nv); R(cacc, 6 * nb); R(cfrc_int, 6 * nb);
  R(efc_J, njm * nv); R(efc_pos, njm); R(efc_margin, njm); R(efc_D, njm); R(efc_vel, njm); R(efc_aref, njm);
  R(efc_frictionloss, njm); R(efc_force, njm); R(efc_Ma, nv);
  R(con_dist, ncm); R(con_pos, 3 * ncm); R(con_frame, 9 * ncm); R(con_includemargin, ncm); R(con_friction, 5 * ncm);
  R(con_solref, 2 * ncm); R(con_solreffriction, 2 * ncm); R(con_solimp, 5 * ncm);
  R(ne, 1); R(nf, 1); R(nl, 1); R(nefc, 1); R(ncon, 1); R(ncollision, 1); R(solver_niter, 1); R(overflow, 1);
  R(efc_type, njm); R(efc_id, njm); R(efc_state, njm); R(moment_rownnz, nu); R(moment_rowadr, nu); R(moment_colind, m->nJmom);
  R(con_dim, ncm); R(con_geom, 2 * ncm); R(con_efc_address, m->nmaxpyramid * ncm); R(con_geomcollisionid, ncm);
  R(eq_active, m->neq);
  R(act, m->na); R(act_dot, m->na); R(ten_length, m->ntendon); R(ten_J, m->nJten); R(ten_velocity, m->ntendon);
  R(cfrc_ext, 6 * nb); R(sensordata, m->nsensordata); R(subtree_linvel, 3 * nb); R(subtree_angmom, 3 * nb);
#undef R
}

/* ------------------------------------------------------------------ kinematics (smooth.py:46-226) */
static void kinematics(W* w) {
  const OrcModel* m = w->m;
  /* world body */
  w->xpos[0] = w->xpos[1] = w->xpos[2] = 0;
  w->xquat[0] = 1; w->xquat[1] = w->xquat[2] = w->xquat[3] = 0;
  for (int b = 1; b < m->nbody; b++) {
    int pid = m->body_parentid[b], jntadr = m->body_jntadr[b], jntnum = m->body_jntnum[b];
    real* xpos = w->xpos + 3 * b; real* xquat = w->xquat + 4 * b;
    if (jntnum == 1 && m->jnt_type[jntadr] == JNT_FREE) { /* smooth.py:84-97 */
      int qa = m->jnt_qposadr[jntadr];
      for (int i = 0; i < 3; i++) xpos[i] = w->qpos[qa + i];
      for (int i = 0; i < 4; i++) xquat[i] = w->qpos[qa + 3 + i];
      normalize4(xquat);
      memcpy(w->xanchor + 3 * jntadr, xpos, 3 * sizeof(real));
      memcpy(w->xaxis + 3 * jntadr, m->jnt_axis + 3 * jntadr, 3 * sizeof(real));
      continue;
    }
    real pos[3], quat[4];
    /* mocap bodies take their pose from Data.mocap_pos / mocap_quat instead of body_pos / body_quat (smooth.py:104-110) */
    int mc = m->body_mocapid[b];
    const real* bpos = mc >= 0 ? w->mocap_pos + 3 * mc : m->body_pos + 3 * b;
    const real* bquat = mc >= 0 ? w->mocap_quat + 4 * mc : m->body_quat + 4 * b;
    rot_vec_quat(bpos, w->xquat + 4 * pid, pos);
    for (int i = 0; i < 3; i++) pos[i] += w->xpos[3 * pid + i];
    mul_quat(w->xquat + 4 * pid, bquat, quat);
    for (int j = jntadr; j < jntadr + jntnum; j++) { /* smooth.py:116-140 */
      int qa = m->jnt_qposadr[j], t = m->jnt_type[j];
      real anchor[3], axis[3], tmp[3];
      rot_vec_quat(m->jnt_pos + 3 * j, quat, anchor);
      for (int i = 0; i < 3; i++) anchor[i] += pos[i];
      rot_vec_quat(m->jnt_axis + 3 * j, quat, axis);
      if (t == JNT_BALL) {
        real ql[4] = {w->qpos[qa], w->qpos[qa + 1], w->qpos[qa + 2], w->qpos[qa + 3]}, nq[4];
        normalize4(ql);
        mul_quat(quat, ql, nq); memcpy(quat, nq, sizeof nq);
        rot_vec_quat(m->jnt_pos + 3 * j, quat, tmp);
        for (int i = 0; i < 3; i++) pos[i] = anchor[i] - tmp[i];
      } else if (t == JNT_SLIDE) {
        for (int i = 0; i < 3; i++) pos[i] += axis[i] * (w->qpos[qa] - m->qpos0[qa]);
      } else if (t == JNT_HINGE) {
        real ql[4], nq[4];
        axis_angle_to_quat(m->jnt_axis + 3 * j, w->qpos[qa] - m->qpos0[qa], ql);
        mul_quat(quat, ql, nq); memcpy(quat, nq, sizeof nq);
        rot_vec_quat(m->jnt_pos + 3 * j, quat, tmp);
        for (int i = 0; i < 3; i++) pos[i] = anchor[i] - tmp[i];
      }
      memcpy(w->xanchor + 3 * j, anchor, sizeof anchor);
      memcpy(w->xaxis + 3 * j, axis, sizeof axis);
    }
    normalize4(quat);
    memcpy(xpos, pos, sizeof pos); memcpy(xquat, quat, sizeof quat);
  }
  for (int b = 0; b < m->nbody; b++) { /* smooth.py:148-175 */
    real t[3], q[4];
    quat_to_mat(w->xquat + 4 * b, w->xmat + 9 * b);
    rot_vec_quat(m->body_ipos + 3 * b, w->xquat + 4 * b, t);
    for (int i = 0; i < 3; i++) w->xipos[3 * b + i] = w->xpos[3 * b + i] + t[i];
    mul_quat(w->xquat + 4 * b, m->body_iquat + 4 * b, q);
    quat_to_mat(q, w->ximat + 9 * b);
  }
  for (int g = 0; g < m->ngeom; g++) { /* smooth.py:178-205: world-welded geoms keep their make_data pose */
    int b = m->geom_bodyid[g];
    if (m->body_weldid[b] == 0 && m->body_mocapid[m->body_rootid[b]] == -1) continue; /* unless it hangs off a mocap body */
    real t[3], q[4];
    rot_vec_quat(m->geom_pos + 3 * g, w->xquat + 4 * b, t);
    for (int i = 0; i < 3; i++) w->geom_xpos[3 * g + i] = w->xpos[3 * b + i] + t[i];
    mul_quat(w->xquat + 4 * b, m->geom_quat + 4 * g, q);
    quat_to_mat(q, w->geom_xmat + 9 * g);
  }
  for (int s = 0; s < m->nsite; s++) { /* smooth.py:208-226 */
    int b = m->site_bodyid[s];
    real t[3], q[4];
    rot_vec_quat(m->site_pos + 3 * s, w->xquat + 4 * b, t);
    for (int i = 0; i < 3; i++) w->site_xpos[3 * s + i] = w->xpos[3 * b + i] + t[i];
    mul_quat(w->xquat + 4 * b, m->site_quat + 4 * s, q);
    quat_to_mat(q, w->site_xmat + 9 * s);
  }
}

/* ------------------------------------------------------------------ com_pos (smooth.py:686-855) */
static void com_pos(W* w) {
  const OrcModel* m = w->m;
  const int nb = m->nbody;
  for (int b = 0; b < nb; b++) for (int i = 0; i < 3; i++) w->subtree_com[3 * b + i] = w->xipos[3 * b + i] * m->body_mass[b];
  for (int b = nb - 1; b >= 1; b--) { int p = m->body_parentid[b]; for (int i = 0; i < 3; i++) w->subtree_com[3 * p + i] += w->subtree_com[3 * b + i]; }
  for (int b = 0; b < nb; b++) { real ms = m->body_subtreemass[b]; if (ms != 0) for (int i = 0; i < 3; i++) w->subtree_com[3 * b + i] /= ms; }
  for (int b = 0; b < nb; b++) { /* _cinert smooth.py:733 */
    const real* mat = w->ximat + 9 * b; const real* inert = m->body_inertia + 3 * b; real mass = m->body_mass[b];
    real dif[3], tmp[9], *res = w->cinert + 10 * b;
    for (int i = 0; i < 3; i++) dif[i] = w->xipos[3 * b + i] - w->subtree_com[3 * m->body_rootid[b] + i];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
      real s = 0; for (int k = 0; k < 3; k++) s += mat[3 * r + k] * inert[k] * mat[3 * c + k]; tmp[3 * r + c] = s;
    }
    res[0] = tmp[0] + mass * (dif[1] * dif[1] + dif[2] * dif[2]);
    res[1] = tmp[4] + mass * (dif[0] * dif[0] + dif[2] * dif[2]);
    res[2] = tmp[8] + mass * (dif[0] * dif[0] + dif[1] * dif[1]);
    res[3] = tmp[1] - mass * dif[0] * dif[1];
    res[4] = tmp[2] - mass * dif[0] * dif[2];
    res[5] = tmp[5] - mass * dif[1] * dif[2];
    res[6] = mass * dif[0]; res[7] = mass * dif[1]; res[8] = mass * dif[2]; res[9] = mass;
  }
  for (int j = 0; j < m->njnt; j++) { /* _cdof smooth.py:779 */
    int b = m->jnt_bodyid[j], d = m->jnt_dofadr[j], t = m->jnt_type[j];
    const real* xaxis = w->xaxis + 3 * j; const real* xm = w->xmat + 9 * b;
    real offset[3], col[3], cr[3];
    for (int i = 0; i < 3; i++) offset[i] = w->subtree_com[3 * m->body_rootid[b] + i] - w->xanchor[3 * j + i];
    real* res = w->cdof + 6 * d;
    if (t == JNT_FREE || t == JNT_BALL) {
      if (t == JNT_FREE) {
        memset(res, 0, 18 * sizeof(real));
        res[3] = 1; res[6 + 4] = 1; res[12 + 5] = 1;
        res += 18;
      }
      for (int k = 0; k < 3; k++) {
        col[0] = xm[k]; col[1] = xm[3 + k]; col[2] = xm[6 + k]; /* column k of xmat = row k of transpose */
        cross3(col, offset, cr);
        for (int i = 0; i < 3; i++) { res[6 * k + i] = col[i]; res[6 * k + 3 + i] = cr[i]; }
      }
    } else if (t == JNT_SLIDE) {
      for (int i = 0; i < 3; i++) { res[i] = 0; res[3 + i] = xaxis[i]; }
    } else {
      cross3(xaxis, offset, cr);
      for (int i = 0; i < 3; i++) { res[i] = xaxis[i]; res[3 + i] = cr[i]; }
    }
  }
}

/* ------------------------------------------------------------------ camlight (smooth.py:858-1027) */
static void camlight(W* w) {
  const OrcModel* m = w->m;
  for (int c = 0; c < m->ncam; c++) {
    int mode = m->cam_mode[c], b = m->cam_bodyid[c], tb = m->cam_targetbodyid[c];
    real* xp = w->cam_xpos + 3 * c; real* xm = w->cam_xmat + 9 * c;
    int is_target = mode == CAM_TARGETBODY || mode == CAM_TARGETBODYCOM;
    real t[3], q[4];
    if (mode == CAM_TRACK && !(is_target && tb < 0)) {
      memcpy(xm, m->cam_mat0 + 9 * c, 9 * sizeof(real));
      for (int i = 0; i < 3; i++) xp[i] = w->xpos[3 * b + i] + m->cam_pos0[3 * c + i];
    } else if (mode == CAM_TRACKCOM) {
      memcpy(xm, m->cam_mat0 + 9 * c, 9 * sizeof(real));
      for (int i = 0; i < 3; i++) xp[i] = w->subtree_com[3 * b + i] + m->cam_poscom0[3 * c + i];
    } else if (is_target && tb >= 0) {
      rot_vec_quat(m->cam_pos + 3 * c, w->xquat + 4 * b, t);
      for (int i = 0; i < 3; i++) xp[i] = w->xpos[3 * b + i] + t[i];
      const real* pos = (mode == CAM_TARGETBODYCOM) ? w->subtree_com + 3 * tb : w->xpos + 3 * tb;
      real m3[3] = {xp[0] - pos[0], xp[1] - pos[1], xp[2] - pos[2]}, m1[3], m2[3], z[3] = {0, 0, 1};
      normalize3(m3); cross3(z, m3, m1); normalize3(m1); cross3(m3, m1, m2); normalize3(m2);
      for (int i = 0; i < 3; i++) { xm[3 * i] = m1[i]; xm[3 * i + 1] = m2[i]; xm[3 * i + 2] = m3[i]; }
    } else {
      rot_vec_quat(m->cam_pos + 3 * c, w->xquat + 4 * b, t);
      for (int i = 0; i < 3; i++) xp[i] = w->xpos[3 * b + i] + t[i];
      mul_quat(w->xquat + 4 * b, m->cam_quat + 4 * c, q);
      quat_to_mat(q, xm);
    }
  }
  for (int l = 0; l < m->nlight; l++) {
    int mode = m->light_mode[l], b = m->light_bodyid[l], tb = m->light_targetbodyid[l];
    real* xp = w->light_xpos + 3 * l; real* xd = w->light_xdir + 3 * l;
    int is_target = mode == CAM_TARGETBODY || mode == CAM_TARGETBODYCOM;
    real t[3];
    if (is_target && tb < 0) { /* invalid target: fixed pose, returns before the normalize (smooth.py:951-958) */
      rot_vec_quat(m->light_pos + 3 * l, w->xquat + 4 * b, t);
      for (int i = 0; i < 3; i++) xp[i] = w->xpos[3 * b + i] + t[i];
      rot_vec_quat(m->light_dir + 3 * l, w->xquat + 4 * b, xd);
      continue;
    } else if (mode == CAM_TRACK) {
      memcpy(xd, m->light_dir0 + 3 * l, 3 * sizeof(real));
      for (int i = 0; i < 3; i++) xp[i] = w->xpos[3 * b + i] + m->light_pos0[3 * l + i];
    } else if (mode == CAM_TRACKCOM) {
      memcpy(xd, m->light_dir0 + 3 * l, 3 * sizeof(real));
      for (int i = 0; i < 3; i++) xp[i] = w->subtree_com[3 * b + i] + m->light_poscom0[3 * l + i];
    } else if (is_target) {
      rot_vec_quat(m->light_pos + 3 * l, w->xquat + 4 * b, t);
      for (int i = 0; i < 3; i++) xp[i] = w->xpos[3 * b + i] + t[i];
      const real* pos = (mode == CAM_TARGETBODYCOM) ? w->subtree_com + 3 * tb : w->xpos + 3 * tb;
      for (int i = 0; i < 3; i++) xd[i] = pos[i] - xp[i];
    } else {
      rot_vec_quat(m->light_pos + 3 * l, w->xquat + 4 * b, t);
      for (int i = 0; i < 3; i++) xp[i] = w->xpos[3 * b + i] + t[i];
      rot_vec_quat(m->light_dir + 3 * l, w->xquat + 4 * b, xd);
    }
    normalize3(xd);
  }
}

/* ------------------------------------------------------------------ crb (smooth.py:1029-1098) */
static void crb(W* w) {
  const OrcModel* m = w->m;
  memcpy(w->crb, w->cinert, 10 * m->nbody * sizeof(real));
  for (int b = m->nbody - 1; b >= 1; b--) {
    int p = m->body_parentid[b];
    if (p == 0) continue;
    for (int i = 0; i < 10; i++) w->crb[10 * p + i] += w->crb[10 * b + i];
  }
  memset(w->M, 0, m->nC * sizeof(real));
  for (int d0 = 0; d0 < m->nv; d0++) { /* _M smooth.py:1048 */
    int b = m->dof_bodyid[d0], madr = m->M_rowadr[d0] + m->M_rownnz[d0] - 1;
    w->M[madr] = m->dof_armature[d0];
    real buf[6];
    inert_vec(w->crb + 10 * b, w->cdof + 6 * d0, buf);
    for (int d = d0; d >= 0; d = m->dof_parentid[d]) {
      real s = 0; for (int i = 0; i < 6; i++) s += w->cdof[6 * d + i] * buf[i];
      w->M[madr] += s; madr--;
    }
  }
}

/* ------------------------------------------------------------------ dense per-tree Cholesky (smooth.py:3227-3265, io.py:173-211)
 * qLD block of tree t: size x size row-major upper factor U (A = U^T U), strictly lower part zero. */
static void tree_dense(const OrcModel* m, const real* Mcsr, int start, int size, real* A /*size*size*/, const real* diag_add) {
  memset(A, 0, (size_t)size * size * sizeof(real));
  for (int i = start; i < start + size; i++) {
    int adr = m->M_rowadr[i];
    for (int k = 0; k < m->M_rownnz[i]; k++) {
      int j = m->M_colind[adr + k];
      real v = Mcsr[adr + k];
      A[(i - start) * size + (j - start)] = v;
      A[(j - start) * size + (i - start)] = v;
    }
    if (diag_add) A[(i - start) * size + (i - start)] += diag_add[i];
  }
}
/* in-place upper Cholesky: returns U in the upper triangle (row-major), zeros below */
static void chol_upper(real* A, int n) {
  for (int j = 0; j < n; j++) {
    real s = A[j * n + j];
    for (int k = 0; k < j; k++) s -= A[k * n + j] * A[k * n + j];
    real ujj = (real)sqrt((double)s);
    A[j * n + j] = ujj;
    for (int i = j + 1; i < n; i++) {
      real t = A[j * n + i];
      for (int k = 0; k < j; k++) t -= A[k * n + j] * A[k * n + i];
      A[j * n + i] = t / ujj;
    }
  }
  for (int i = 1; i < n; i++) for (int j = 0; j < i; j++) A[i * n + j] = 0;
}
static void chol_upper_solve(const real* U, int n, const real* b, real* x) {
  /* U^T y = b ; U x = y */
  for (int i = 0; i < n; i++) { real s = b[i]; for (int k = 0; k < i; k++) s -= U[k * n + i] * x[k]; x[i] = s / U[i * n + i]; }
  for (int i = n - 1; i >= 0; i--) { real s = x[i]; for (int k = i + 1; k < n; k++) s -= U[i * n + k] * x[k]; x[i] = s / U[i * n + i]; }
}
/* factor_solve_i (smooth.py:3352): factor M (+diag_add) into qLD-like storage L, solve x = (M+diag)^-1 y */
static void factor_solve_i(W* w, const real* Mcsr, const real* diag_add, real* L, real* x, const real* y) {
  const OrcModel* m = w->m;
  for (int t = 0; t < m->ntree; t++) {
    int start = m->tree_dofadr[t], size = m->tree_dofnum[t];
    real* blk = L + m->qLD_block_adr[start];
    tree_dense(m, Mcsr, start, size, blk, diag_add);
    chol_upper(blk, size);
    chol_upper_solve(blk, size, y + start, x + start);
  }
}
/* mul_m (support.py:153): res = M vec using the CSR lower triangle */
static void mul_m(const OrcModel* m, const real* Mcsr, const real* vec, real* res) {
  for (int i = 0; i < m->nv; i++) res[i] = 0;
  for (int i = 0; i < m->nv; i++) {
    int adr = m->M_rowadr[i], n = m->M_rownnz[i];
    for (int k = 0; k < n; k++) {
      int j = m->M_colind[adr + k];
      res[i] += Mcsr[adr + k] * vec[j];
      if (j != i) res[j] += Mcsr[adr + k] * vec[i];
    }
  }
}

/* ------------------------------------------------------------------ tendon (smooth.py:3658-3692, 4197; fixed tendons: joint wraps) */
static void tendon(W* w) {
  const OrcModel* m = w->m;
  for (int t = 0; t < m->ntendon; t++) w->ten_length[t] = 0;
  for (int i = 0; i < m->nJten; i++) w->ten_J[i] = 0;
  for (int t = 0; t < m->ntendon; t++)
    for (int k = m->tendon_adr[t]; k < m->tendon_adr[t] + m->tendon_num[t]; k++) {
      const int j = m->wrap_objid[k], dof = m->jnt_dofadr[j];
      const real prm = m->wrap_prm[k];
      w->ten_length[t] += prm * w->qpos[m->jnt_qposadr[j]];
      for (int c = 0; c < m->ten_J_rownnz[t]; c++)
        if (m->ten_J_colind[m->ten_J_rowadr[t] + c] == dof) { w->ten_J[m->ten_J_rowadr[t] + c] = prm; break; }
    }
}
/* dense row (nv) of tendon t's Jacobian times scale, added into out */
static void tendon_row(const W* w, int t, real scale, real* out) {
  const OrcModel* m = w->m;
  for (int c = 0; c < m->ten_J_rownnz[t]; c++) { const int s = m->ten_J_rowadr[t] + c; out[m->ten_J_colind[s]] += scale * w->ten_J[s]; }
}

/* ------------------------------------------------------------------ transmission (smooth.py:2288-2396, :2508-2525; joint and tendon transmission) */
static void transmission(W* w) {
  const OrcModel* m = w->m;
  int nnz = 0;
  for (int a = 0; a < m->nu; a++) {
    if (m->actuator_trntype[a] == TRN_TENDON) {
      const int t = m->actuator_trnid[2 * a];
      const real gear0 = m->actuator_gear[6 * a];
      w->actuator_length[a] = w->ten_length[t] * gear0;
      w->moment_rownnz[a] = m->ten_J_rownnz[t]; w->moment_rowadr[a] = nnz;
      for (int c = 0; c < m->ten_J_rownnz[t]; c++) { w->moment_colind[nnz] = m->ten_J_colind[m->ten_J_rowadr[t] + c]; w->actuator_moment[nnz] = w->ten_J[m->ten_J_rowadr[t] + c] * gear0; nnz++; }
      continue;
    }
    int j = m->actuator_trnid[2 * a], t = m->jnt_type[j], qa = m->jnt_qposadr[j], va = m->jnt_dofadr[j];
    const real* gear = m->actuator_gear + 6 * a;
    if (t == JNT_SLIDE || t == JNT_HINGE) {
      w->actuator_length[a] = w->qpos[qa] * gear[0];
      w->moment_rownnz[a] = 1; w->moment_rowadr[a] = nnz;
      w->moment_colind[nnz] = va; w->actuator_moment[nnz] = gear[0];
      nnz += 1;
    } else if (t == JNT_FREE) {
      w->actuator_length[a] = 0;
      w->moment_rownnz[a] = 6; w->moment_rowadr[a] = nnz;
      for (int i = 0; i < 6; i++) { w->moment_colind[nnz + i] = va + i; w->actuator_moment[nnz + i] = gear[i]; }
      nnz += 6;
    } else {
      w->overflow[0] |= OVF_UNSUPPORTED; /* ball-joint transmission not restated */
      w->moment_rownnz[a] = 0; w->moment_rowadr[a] = nnz; w->actuator_length[a] = 0;
    }
  }
}

/* ------------------------------------------------------------------ collision */
/* collision_driver.py:98-113 */
static int plane_filter(real size1, real size2, real margin1, real margin2, const real* xpos1, const real* xpos2, const real* xmat1, const real* xmat2) {
  if (size1 == 0) {
    real n[3] = {xmat1[2], xmat1[5], xmat1[8]}, d[3] = {xpos2[0] - xpos1[0], xpos2[1] - xpos1[1], xpos2[2] - xpos1[2]};
    return dot3(d, n) <= size2 + margin1 + margin2;
  } else if (size2 == 0) {
    real n[3] = {xmat2[2], xmat2[5], xmat2[8]}, d[3] = {xpos1[0] - xpos2[0], xpos1[1] - xpos2[1], xpos1[2] - xpos2[2]};
    return dot3(d, n) <= size1 + margin1 + margin2;
  }
  return 1;
}
/* collision_driver.py:116-121 */
static int sphere_filter(real size1, real size2, real margin1, real margin2, const real* xpos1, const real* xpos2) {
  real bound = size1 + size2 + margin1 + margin2, dif[3] = {xpos2[0] - xpos1[0], xpos2[1] - xpos1[1], xpos2[2] - xpos1[2]};
  return dot3(dif, dif) <= bound * bound;
}
static void matvec3(const real* m, const real* v, real* o) {
  real r0 = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], r1 = m[3] * v[0] + m[4] * v[1] + m[5] * v[2], r2 = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  o[0] = r0; o[1] = r1; o[2] = r2;
}
/* collision_driver.py:124-219 */
static int aabb_filter(const real* c1, const real* c2, const real* s1, const real* s2, real margin1, real margin2, const real* xpos1, const real* xpos2, const real* xmat1, const real* xmat2) {
  real ce1[3], ce2[3], margin = margin1 + margin2;
  matvec3(xmat1, c1, ce1); matvec3(xmat2, c2, ce2);
  for (int i = 0; i < 3; i++) { ce1[i] += xpos1[i]; ce2[i] += xpos2[i]; }
  real mx1[3] = {-MJ_MAXVAL, -MJ_MAXVAL, -MJ_MAXVAL}, mn1[3] = {MJ_MAXVAL, MJ_MAXVAL, MJ_MAXVAL};
  real mx2[3] = {-MJ_MAXVAL, -MJ_MAXVAL, -MJ_MAXVAL}, mn2[3] = {MJ_MAXVAL, MJ_MAXVAL, MJ_MAXVAL};
  real sg[2] = {-1, 1};
  for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
    real co1[3] = {sg[i] * s1[0], sg[j] * s1[1], sg[k] * s1[2]}, co2[3] = {sg[i] * s2[0], sg[j] * s2[1], sg[k] * s2[2]}, p1[3], p2[3];
    matvec3(xmat1, co1, p1); matvec3(xmat2, co2, p2);
    for (int a = 0; a < 3; a++) {
      if (p1[a] > mx1[a]) mx1[a] = p1[a];
      if (p1[a] < mn1[a]) mn1[a] = p1[a];
      if (p2[a] > mx2[a]) mx2[a] = p2[a];
      if (p2[a] < mn2[a]) mn2[a] = p2[a];
    }
  }
  for (int a = 0; a < 3; a++) {
    if (ce1[a] + mx1[a] + margin < ce2[a] + mn2[a]) return 0;
    if (ce2[a] + mx2[a] + margin < ce1[a] + mn1[a]) return 0;
  }
  return 1;
}
/* collision_driver.py:223-279 */
static int obb_filter(const real* c1, const real* c2, const real* s1, const real* s2, real margin1, real margin2, const real* xpos1, const real* xpos2, const real* xmat1, const real* xmat2) {
  real margin = margin1 + margin2, xc[2][3], normal[6][3];
  matvec3(xmat1, c1, xc[0]); matvec3(xmat2, c2, xc[1]);
  for (int i = 0; i < 3; i++) { xc[0][i] += xpos1[i]; xc[1][i] += xpos2[i]; }
  for (int k = 0; k < 3; k++) for (int i = 0; i < 3; i++) { normal[k][i] = xmat1[3 * i + k]; normal[3 + k][i] = xmat2[3 * i + k]; }
  for (int j = 0; j < 2; j++) for (int k = 0; k < 3; k++) {
    real proj[2], radius[2];
    for (int i = 0; i < 2; i++) {
      const real* size = i == 0 ? s1 : s2;
      proj[i] = dot3(xc[i], normal[3 * j + k]);
      radius[i] = (real)fabs((double)(size[0] * dot3(normal[3 * i + 0], normal[3 * j + k]))) + (real)fabs((double)(size[1] * dot3(normal[3 * i + 1], normal[3 * j + k]))) +
                  (real)fabs((double)(size[2] * dot3(normal[3 * i + 2], normal[3 * j + k])));
    }
    if (radius[0] + radius[1] + margin < (real)fabs((double)(proj[1] - proj[0]))) return 0;
  }
  return 1;
}
/* collision_driver.py:282-334 */
static int broadphase_filter(const W* w, int g1, int g2) {
  const OrcModel* m = w->m;
  real rb1 = m->geom_rbound[g1], rb2 = m->geom_rbound[g2];
  real em1 = m->geom_margin[g1] + m->geom_gap[g1], em2 = m->geom_margin[g2] + m->geom_gap[g2];
  const real *xp1 = w->geom_xpos + 3 * g1, *xp2 = w->geom_xpos + 3 * g2, *xm1 = w->geom_xmat + 9 * g1, *xm2 = w->geom_xmat + 9 * g2;
  int f = m->broadphase_filter;
  if (rb1 == 0 || rb2 == 0) {
    if (f & BF_PLANE) return plane_filter(rb1, rb2, em1, em2, xp1, xp2, xm1, xm2);
  } else {
    const real *c1 = m->geom_aabb + 6 * g1, *c2 = m->geom_aabb + 6 * g2;
    if ((f & BF_SPHERE) && !sphere_filter(rb1, rb2, em1, em2, xp1, xp2)) return 0;
    if ((f & BF_AABB) && !aabb_filter(c1, c2, c1 + 3, c2 + 3, em1, em2, xp1, xp2, xm1, xm2)) return 0;
    if ((f & BF_OBB) && !obb_filter(c1, c2, c1 + 3, c2 + 3, em1, em2, xp1, xp2, xm1, xm2)) return 0;
  }
  return 1;
}

typedef struct { real margin, gap; int condim; real friction[5], solref[2], solreffriction[2], solimp[5]; } ConParams;
/* collision_core.py:294-412 (geom pairs only: pairid == -1; adhesion is zero on this path) */
static void contact_params(const OrcModel* m, int g1, int g2, int pairid, ConParams* p) {
  if (pairid > -1) { /* explicit <pair>: parameters come from the pair (collision_core.py:305-307, 343-349) */
    p->margin = m->pair_margin[pairid]; p->gap = m->pair_gap[pairid]; p->condim = m->pair_dim[pairid];
    for (int i = 0; i < 5; i++) { p->friction[i] = rmax(MJ_MINMU, m->pair_friction[5 * pairid + i]); p->solimp[i] = m->pair_solimp[5 * pairid + i]; }
    for (int i = 0; i < 2; i++) { p->solref[i] = m->pair_solref[2 * pairid + i]; p->solreffriction[i] = m->pair_solreffriction[2 * pairid + i]; }
    return;
  }
  p->margin = m->geom_margin[g1] + m->geom_margin[g2];
  p->gap = m->geom_gap[g1] + m->geom_gap[g2];
  real solmix1 = m->geom_solmix[g1], solmix2 = m->geom_solmix[g2], mix;
  int p1 = m->geom_priority[g1], p2 = m->geom_priority[g2];
  real fr[3];
  if (p1 > p2) { mix = 1; p->condim = m->geom_condim[g1]; memcpy(fr, m->geom_friction + 3 * g1, sizeof fr); }
  else if (p2 > p1) { mix = 0; p->condim = m->geom_condim[g2]; memcpy(fr, m->geom_friction + 3 * g2, sizeof fr); }
  else {
    mix = safe_div(solmix1, solmix1 + solmix2);
    if (solmix1 < MJ_MINVAL && solmix2 < MJ_MINVAL) mix = (real)0.5;
    if (solmix1 < MJ_MINVAL && solmix2 >= MJ_MINVAL) mix = 0;
    if (solmix1 >= MJ_MINVAL && solmix2 < MJ_MINVAL) mix = 1;
    p->condim = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
    for (int i = 0; i < 3; i++) fr[i] = rmax(m->geom_friction[3 * g1 + i], m->geom_friction[3 * g2 + i]);
  }
  p->friction[0] = fr[0]; p->friction[1] = fr[0]; p->friction[2] = fr[1]; p->friction[3] = fr[2]; p->friction[4] = fr[2];
  const real *sr1 = m->geom_solref + 2 * g1, *sr2 = m->geom_solref + 2 * g2;
  if (sr1[0] > 0 && sr2[0] > 0) for (int i = 0; i < 2; i++) p->solref[i] = mix * sr1[i] + (1 - mix) * sr2[i];
  else for (int i = 0; i < 2; i++) p->solref[i] = rmin(sr1[i], sr2[i]);
  p->solreffriction[0] = p->solreffriction[1] = 0;
  for (int i = 0; i < 5; i++) p->solimp[i] = mix * m->geom_solimp[5 * g1 + i] + (1 - mix) * m->geom_solimp[5 * g2 + i];
  for (int i = 0; i < 5; i++) p->friction[i] = rmax(MJ_MINMU, p->friction[i]);
}
/* collision_core.py:213-291 (per-world slots; pairid[0] == -1, pairid[1] == -1 on this path) */
static void write_contact(W* w, int id, real dist, const real* pos, const real* frame, const ConParams* p, int g1, int g2) {
  int detected = dist < p->margin + p->gap;
  if (!detected) return;
  int cid = w->ncon[0]++;
  if (cid >= w->nconmax) { w->overflow[0] |= OVF_NARROWPHASE; return; }
  const int np = w->m->nmaxpyramid;
  w->con_dist[cid] = dist;
  memcpy(w->con_pos + 3 * cid, pos, 3 * sizeof(real));
  memcpy(w->con_frame + 9 * cid, frame, 9 * sizeof(real));
  w->con_geom[2 * cid] = g1; w->con_geom[2 * cid + 1] = g2;
  w->con_includemargin[cid] = p->margin;
  w->con_dim[cid] = p->condim;
  memcpy(w->con_friction + 5 * cid, p->friction, 5 * sizeof(real));
  memcpy(w->con_solref + 2 * cid, p->solref, 2 * sizeof(real));
  memcpy(w->con_solreffriction + 2 * cid, p->solreffriction, 2 * sizeof(real));
  memcpy(w->con_solimp + 5 * cid, p->solimp, 5 * sizeof(real));
  w->con_geomcollisionid[cid] = id;
  for (int i = 0; i < np; i++) w->con_efc_address[np * cid + i] = -1;
}
/* collision_primitive_core.py:47 */
static real plane_sphere(const real* n, const real* ppos, const real* spos, real r, real* pos) {
  real d[3] = {spos[0] - ppos[0], spos[1] - ppos[1], spos[2] - ppos[2]};
  real dist = dot3(d, n) - r;
  for (int i = 0; i < 3; i++) pos[i] = spos[i] - n[i] * (r + (real)0.5 * dist);
  return dist;
}
/* collision_primitive_core.py:55 */
static real sphere_sphere(const real* pos1, real r1, const real* pos2, real r2, real* pos, real* n) {
  real dir[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
  real dist = len3(dir);
  if (dist == 0) { n[0] = 1; n[1] = 0; n[2] = 0; } else for (int i = 0; i < 3; i++) n[i] = dir[i] / dist;
  dist = dist - (r1 + r2);
  for (int i = 0; i < 3; i++) pos[i] = pos1[i] + n[i] * (r1 + (real)0.5 * dist);
  return dist;
}
/* ------------------------------------------------------------------ box / cylinder / ellipsoid primitives
 * (collision_primitive_core.py:305-1433).  Outputs use MAXVAL for unpopulated contact slots like the reference. */
#define ORC_MAXVAL ((real)1e10)
static inline void matT_vec3(const real* m, const real* v, real* o) {
  for (int j = 0; j < 3; j++) o[j] = m[j] * v[0] + m[3 + j] * v[1] + m[6 + j] * v[2];
}
static inline void mat_mul3(const real* a, const real* b, real* c) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
static inline void mat_T3(const real* a, real* t) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * j + i]; }
static inline real rabs(real x) { return x < 0 ? -x : x; }
/* collision_primitive_core.py:305 */
static real plane_ellipsoid(const real* n, const real* ppos, const real* epos, const real* erot, const real* esize, real* pos) {
  real loc[3], sup[3], off[3];
  matT_vec3(erot, n, loc);
  for (int i = 0; i < 3; i++) loc[i] *= esize[i];
  normalize3(loc);
  for (int i = 0; i < 3; i++) sup[i] = -loc[i] * esize[i];
  matvec3(erot, sup, off);
  for (int i = 0; i < 3; i++) pos[i] = epos[i] + off[i];
  real d[3] = {pos[0] - ppos[0], pos[1] - ppos[1], pos[2] - ppos[2]};
  real dist = dot3(n, d);
  for (int i = 0; i < 3; i++) pos[i] -= n[i] * dist * (real)0.5;
  return dist;
}
/* collision_primitive_core.py:336: all 8 corners, the caller keeps those within margin */
static void plane_box(const real* n, const real* ppos, const real* bpos, const real* brot, const real* bsize, real dist[8], real pos[8][3]) {
  real d[3] = {bpos[0] - ppos[0], bpos[1] - ppos[1], bpos[2] - ppos[2]};
  real center_dist = dot3(d, n);
  for (int i = 0; i < 8; i++) {
    real cl[3] = {(i & 1) ? bsize[0] : -bsize[0], (i & 2) ? bsize[1] : -bsize[1], (i & 4) ? bsize[2] : -bsize[2]}, c[3];
    matvec3(brot, cl, c);
    real cdist = center_dist + dot3(n, c);
    dist[i] = cdist;
    for (int k = 0; k < 3; k++) pos[i][k] = c[k] + bpos[k] - (real)0.5 * n[k] * cdist;
  }
}
/* collision_primitive_core.py:387 */
static real sphere_cylinder(const real* spos, real sr, const real* cpos, const real* caxis, real cr, real chh, real* pos, real* nrm) {
  real vec[3] = {spos[0] - cpos[0], spos[1] - cpos[1], spos[2] - cpos[2]};
  real x = dot3(vec, caxis), a_proj[3], p_proj[3];
  for (int i = 0; i < 3; i++) { a_proj[i] = caxis[i] * x; p_proj[i] = vec[i] - a_proj[i]; }
  real p_proj_sqr = dot3(p_proj, p_proj);
  int collide_side = rabs(x) < chh, collide_cap = p_proj_sqr < cr * cr;
  if (collide_side && collide_cap) {
    real dist_cap = chh - rabs(x), dist_radius = cr - (real)sqrt((double)p_proj_sqr);
    if (dist_cap < dist_radius) collide_side = 0; else collide_cap = 0;
  }
  if (collide_side) {
    real tgt[3] = {cpos[0] + a_proj[0], cpos[1] + a_proj[1], cpos[2] + a_proj[2]};
    return sphere_sphere(spos, sr, tgt, cr, pos, nrm);
  } else if (collide_cap) {
    real sgn = x > 0 ? (real)1 : (real)-1, pcap[3], pn[3];
    for (int i = 0; i < 3; i++) { pcap[i] = cpos[i] + sgn * caxis[i] * chh; pn[i] = sgn * caxis[i]; }
    real dist = plane_sphere(pn, pcap, spos, sr, pos);
    for (int i = 0; i < 3; i++) nrm[i] = -pn[i];
    return dist;
  }
  real inv_len = safe_div(1, (real)sqrt((double)p_proj_sqr)), sgn = x < 0 ? (real)-1 : (real)1 /* wp.sign(0) = +1 */, corner[3];
  for (int i = 0; i < 3; i++) corner[i] = cpos[i] + caxis[i] * (sgn * chh) + p_proj[i] * (cr * inv_len);
  return sphere_sphere(spos, sr, corner, 0, pos, nrm);
}
/* collision_primitive_core.py:459: two rim points + a triangle on the near cap */
static void plane_cylinder(const real* n, const real* ppos, const real* center, const real* caxis, real cr, real chh, real dist[4], real pos[4][3]) {
  real axis[3] = {caxis[0], caxis[1], caxis[2]};
  real prjaxis = dot3(n, axis);
  if (prjaxis > 0) { for (int i = 0; i < 3; i++) axis[i] = -axis[i]; prjaxis = -prjaxis; }
  real d0v[3] = {center[0] - ppos[0], center[1] - ppos[1], center[2] - ppos[2]};
  real dist0 = dot3(d0v, n), vec[3];
  for (int i = 0; i < 3; i++) vec[i] = axis[i] * prjaxis - n[i];
  real len_sqr = dot3(vec, vec);
  if (len_sqr >= (real)1e-12) { real sc = safe_div(cr, (real)sqrt((double)len_sqr)); for (int i = 0; i < 3; i++) vec[i] *= sc; }
  else { vec[0] = cr; vec[1] = 0; vec[2] = 0; }
  real prjvec = dot3(vec, n);
  for (int i = 0; i < 3; i++) axis[i] *= chh;
  prjaxis *= chh;
  real dist1 = dist0 + prjaxis + prjvec, dist2 = dist0 - prjaxis + prjvec;
  for (int i = 0; i < 3; i++) { pos[0][i] = center[i] + vec[i] + axis[i] - n[i] * (dist1 * (real)0.5); pos[1][i] = center[i] + vec[i] - axis[i] - n[i] * (dist2 * (real)0.5); }
  dist[0] = dist1; dist[1] = dist2;
  real prjvec1 = -prjvec * (real)0.5, dist3 = dist0 + prjaxis + prjvec1, vec1[3];
  cross3(vec, axis, vec1);
  normalize3(vec1);
  real sc = cr * (real)sqrt(3.0) * (real)0.5;
  for (int i = 0; i < 3; i++) vec1[i] *= sc;
  for (int i = 0; i < 3; i++) {
    pos[2][i] = center[i] + vec1[i] + axis[i] - vec[i] * (real)0.5 - n[i] * (dist3 * (real)0.5);
    pos[3][i] = center[i] - vec1[i] + axis[i] - vec[i] * (real)0.5 - n[i] * (dist3 * (real)0.5);
  }
  dist[2] = dist3; dist[3] = dist3;
}
/* collision_primitive_core.py:1043 */
static real sphere_box(const real* spos, real sr, const real* bpos, const real* brot, const real* bsize, real* cpos, real* nrm) {
  real rel[3] = {spos[0] - bpos[0], spos[1] - bpos[1], spos[2] - bpos[2]}, center[3], clamped[3], dir[3], pos[3], cdist;
  matT_vec3(brot, rel, center);
  for (int i = 0; i < 3; i++) { clamped[i] = rmax(-bsize[i], rmin(bsize[i], center[i])); dir[i] = clamped[i] - center[i]; }
  real dist = len3(dir);
  if (dist <= MJ_MINVAL) { /* centre inside the box: push out through the nearest face */
    real closest = 2 * (bsize[0] + bsize[1] + bsize[2]); int k = 0;
    for (int i = 0; i < 6; i++) {
      real face_dist = rabs(((i % 2) ? (real)1 : (real)-1) * bsize[i / 2] - center[i / 2]);
      if (closest > face_dist) { closest = face_dist; k = i; }
    }
    real nearest[3] = {0, 0, 0};
    nearest[k / 2] = (k % 2) ? (real)-1 : (real)1;
    for (int i = 0; i < 3; i++) pos[i] = center[i] + nearest[i] * (sr - closest) / 2;
    matvec3(brot, nearest, nrm);
    cdist = -closest - sr;
  } else {
    for (int i = 0; i < 3; i++) dir[i] /= dist;
    for (int i = 0; i < 3; i++) pos[i] = (real)0.5 * (clamped[i] + center[i] + dir[i] * sr);
    matvec3(brot, dir, nrm);
    cdist = dist - sr;
  }
  real g[3]; matvec3(brot, pos, g);
  for (int i = 0; i < 3; i++) cpos[i] = bpos[i] + g[i];
  return cdist;
}
/* collision_primitive_core.py:1098 (after MuJoCo's mjc_CapsuleBox): closest feature search, then an optional second
 * contact further along the capsule; both contacts are sphere-box tests at the chosen segment points */
static void capsule_box(const real* cpos_in, const real* caxis, real crad, real chl, const real* bpos, const real* brot, const real* bsize,
                        real dist[2], real cpos[2][3], real cnrm[2][3]) {
  real rel[3] = {cpos_in[0] - bpos[0], cpos_in[1] - bpos[1], cpos_in[2] - bpos[2]}, pos[3], axis[3], halfaxis[3];
  matT_vec3(brot, rel, pos);
  matT_vec3(brot, caxis, axis);
  for (int i = 0; i < 3; i++) halfaxis[i] = axis[i] * chl;
  int axisdir = (halfaxis[0] > 0) + 2 * (halfaxis[1] > 0) + 4 * (halfaxis[2] > 0);
  real bestdist = (real)1e32, bestsegmentpos = -12;
  int cltype = -4, clface = -12;
  for (int i = -1; i <= 1; i += 2) { /* a capsule tip closest to a face */
    real tip[3], bp[3]; int n_out = 0, ax_out = -1;
    for (int j = 0; j < 3; j++) { tip[j] = pos[j] + (real)i * halfaxis[j]; bp[j] = tip[j]; }
    for (int j = 0; j < 3; j++) {
      if (bp[j] < -bsize[j]) { n_out++; ax_out = j; bp[j] = -bsize[j]; }
      else if (bp[j] > bsize[j]) { n_out++; ax_out = j; bp[j] = bsize[j]; }
    }
    if (n_out > 1) continue;
    real dd[3] = {bp[0] - tip[0], bp[1] - tip[1], bp[2] - tip[2]}, ds = dot3(dd, dd);
    if (ds < bestdist) { bestdist = ds; bestsegmentpos = (real)i; cltype = -2 + i; clface = ax_out; }
  }
  int clcorner = -123, cledge = -123; real bestboxpos = 0;
  for (int i = 0; i < 8; i++) for (int j = 0; j < 3; j++) { /* box edges (corner i, direction j) */
    if (i & (1 << j)) continue;
    real box_pt[3] = {((i & 1) ? 1 : -1) * bsize[0], ((i & 2) ? 1 : -1) * bsize[1], ((i & 4) ? 1 : -1) * bsize[2]}, dif[3];
    box_pt[j] = 0;
    for (int k = 0; k < 3; k++) dif[k] = box_pt[k] - pos[k];
    real u = -bsize[j] * dif[j], v = dot3(halfaxis, dif), ma = bsize[j] * bsize[j], mb = -bsize[j] * halfaxis[j], mc = chl * chl;
    real det = ma * mc - mb * mb;
    if (rabs(det) < MJ_MINVAL) continue;
    real idet = 1 / det, x1 = (mc * u - mb * v) * idet, x2 = (ma * v - mb * u) * idet;
    int s1 = 1, s2 = 1;
    if (x1 > 1) { x1 = 1; s1 = 2; x2 = safe_div(v - mb, mc); }
    else if (x1 < -1) { x1 = -1; s1 = 0; x2 = safe_div(v + mb, mc); }
    int x2_over = x2 > 1;
    if (x2_over || x2 < -1) {
      if (x2_over) { x2 = 1; s2 = 2; x1 = safe_div(u - mb, ma); } else { x2 = -1; s2 = 0; x1 = safe_div(u + mb, ma); }
      if (x1 > 1) { x1 = 1; s1 = 2; } else if (x1 < -1) { x1 = -1; s1 = 0; }
    }
    for (int k = 0; k < 3; k++) dif[k] -= halfaxis[k] * x2;
    dif[j] += bsize[j] * x1;
    int ct = s1 * 3 + s2;
    real dsq = dot3(dif, dif);
    if (dsq < bestdist - MJ_MINVAL) {
      bestdist = dsq; bestsegmentpos = x2; bestboxpos = x1;
      int c2 = ct / 6;
      clcorner = i + (1 << j) * c2; cledge = j; cltype = ct;
    }
  }
  dist[0] = dist[1] = ORC_MAXVAL;
  memset(cpos, 0, 6 * sizeof(real)); memset(cnrm, 0, 6 * sizeof(real));
  if (cltype == -4) return;
  real secondpos = -4;
  if (cltype >= 0 && cltype / 3 != 1) { /* a box corner is closest */
    int c1 = axisdir ^ clcorner;
    if (c1 != 0 && c1 != 7) {
      int mul, ax = 0, ax1 = 1, ax2 = 2;
      if (c1 == 1 || c1 == 2 || c1 == 4) mul = 1; else { mul = -1; c1 = 7 - c1; }
      if (c1 == 1) { ax = 0; ax1 = 1; ax2 = 2; } else if (c1 == 2) { ax = 1; ax1 = 2; ax2 = 0; } else if (c1 == 4) { ax = 2; ax1 = 0; ax2 = 1; }
      if (axis[ax] * axis[ax] > (real)0.5) {
        real mm = 2 * safe_div(bsize[ax], rabs(halfaxis[ax]));
        secondpos = rmin(1 - (real)mul * bestsegmentpos, mm);
      } else {
        real mm = 2 * rmin(safe_div(bsize[ax1], rabs(halfaxis[ax1])), safe_div(bsize[ax2], rabs(halfaxis[ax2])));
        secondpos = -rmin(1 + (real)mul * bestsegmentpos, mm);
      }
      secondpos *= (real)mul;
    }
  } else if (cltype >= 0 && cltype / 3 == 1) { /* the interior of a box edge is closest */
    int c1 = axisdir ^ clcorner;
    c1 &= 7 - (1 << cledge);
    if (c1 == 1 || c1 == 2 || c1 == 4) {
      int ax = cledge, ax1 = (cledge + 1) % 3, ax2 = (cledge + 2) % 3, mul;
      if (rabs(axis[ax1]) > rabs(axis[ax2])) ax1 = ax2;
      ax2 = 3 - ax - ax1;
      if (c1 & (1 << ax2)) { mul = 1; secondpos = 1 - bestsegmentpos; } else { mul = -1; secondpos = 1 + bestsegmentpos; }
      real e1 = 2 * safe_div(bsize[ax2], rabs(halfaxis[ax2])), e2;
      secondpos = rmin(e1, secondpos);
      if (((axisdir & (1 << ax)) != 0) == ((c1 & (1 << ax2)) != 0)) e2 = 1 - bestboxpos; else e2 = 1 + bestboxpos;
      e1 = bsize[ax] * safe_div(e2, rabs(halfaxis[ax]));
      secondpos = rmin(e1, secondpos);
      secondpos *= (real)mul;
    }
  } else if (cltype < 0) { /* a capsule tip over a face: the second point is the far end clamped to the face */
    if (clface != -1) {
      int mul = cltype == -3 ? 1 : -1;
      secondpos = 2;
      real tmp1[3];
      for (int i = 0; i < 3; i++) tmp1[i] = pos[i] - halfaxis[i] * (real)mul;
      for (int i = 0; i < 3; i++) {
        if (i == clface) continue;
        real ha_r = safe_div((real)mul, halfaxis[i]);
        real e1 = (bsize[i] - tmp1[i]) * ha_r;
        if (0 < e1 && e1 < secondpos) secondpos = e1;
        e1 = (-bsize[i] - tmp1[i]) * ha_r;
        if (0 < e1 && e1 < secondpos) secondpos = e1;
      }
      secondpos *= (real)mul;
    }
  }
  real sl[3], sg[3];
  for (int i = 0; i < 3; i++) sl[i] = pos[i] + halfaxis[i] * bestsegmentpos;
  matvec3(brot, sl, sg);
  for (int i = 0; i < 3; i++) sg[i] += bpos[i];
  dist[0] = sphere_box(sg, crad, bpos, brot, bsize, cpos[0], cnrm[0]);
  if (secondpos > -3) {
    for (int i = 0; i < 3; i++) sl[i] = pos[i] + halfaxis[i] * (secondpos + bestsegmentpos);
    matvec3(brot, sl, sg);
    for (int i = 0; i < 3; i++) sg[i] += bpos[i];
    dist[1] = sphere_box(sg, crad, bpos, brot, bsize, cpos[1], cnrm[1]);
  }
}
/* collision_primitive_core.py:556 */
static void rotmore_of(int face, real* r) {
  for (int i = 0; i < 9; i++) r[i] = 0;
  switch (face) {
    case 0: r[2] = -1; r[4] = 1; r[6] = 1; break;
    case 1: r[0] = 1; r[5] = -1; r[7] = 1; break;
    case 2: r[0] = 1; r[4] = 1; r[8] = 1; break;
    case 3: r[2] = 1; r[4] = 1; r[6] = -1; break;
    case 4: r[0] = 1; r[5] = 1; r[7] = -1; break;
    case 5: r[0] = -1; r[4] = 1; r[8] = -1; break;
  }
}
/* collision_primitive_core.py:588 (after MuJoCo's mjc_BoxBox): separating-axis search over 6 face normals and 9 edge
 * cross products, then clipping of the incident face / edge against the reference face.  Returns the contact count. */
static int box_box(const real* pos1, const real* rot1, const real* size1, const real* pos2, const real* rot2, const real* size2, real margin,
                   real cdist[8], real cpos[8][3], real cnormal[3]) {
  real d21[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]}, d12[3] = {-d21[0], -d21[1], -d21[2]};
  real pos21[3], pos12[3], rot1T[9], rot21[9], rot12[9], rot21abs[9], rot12abs[9], plen1[3], plen2[3];
  matT_vec3(rot1, d21, pos21);
  matT_vec3(rot2, d12, pos12);
  mat_T3(rot1, rot1T);
  mat_mul3(rot1T, rot2, rot21);
  mat_T3(rot21, rot12);
  for (int i = 0; i < 9; i++) rot21abs[i] = rabs(rot21[i]);
  mat_T3(rot21abs, rot12abs);
  matvec3(rot21abs, size2, plen2);
  matvec3(rot12abs, size1, plen1);
  real separation = margin + 3 * (size1[0] + size2[0]) + 3 * (size1[1] + size2[1]) + 3 * (size1[2] + size2[2]);
  int axis_code = -1;
  for (int i = 0; i < 3; i++) {
    real c1 = -rabs(pos21[i]) + size1[i] + plen2[i], c2 = -rabs(pos12[i]) + size2[i] + plen1[i];
    if (c1 < -margin || c2 < -margin) return 0;
    if (c1 < separation) { separation = c1; axis_code = i + 3 * (pos21[i] < 0); }
    if (c2 < separation) { separation = c2; axis_code = i + 3 * (pos12[i] < 0) + 6; }
  }
  real clnorm[3] = {0, 0, 0}; int inv = 0, cle1 = 0, cle2 = 0;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    const real* a = rot12 + 3 * j; real ca[3];
    if (i == 0) { ca[0] = 0; ca[1] = -a[2]; ca[2] = a[1]; } else if (i == 1) { ca[0] = a[2]; ca[1] = 0; ca[2] = -a[0]; } else { ca[0] = -a[1]; ca[1] = a[0]; ca[2] = 0; }
    real cl = len3(ca);
    if (cl < MJ_MINVAL) continue;
    for (int k = 0; k < 3; k++) ca[k] /= cl;
    real box_dist = dot3(pos21, ca), c3 = 0;
    for (int k = 0; k < 3; k++) {
      if (k != i) c3 += size1[k] * rabs(ca[k]);
      if (k != j) c3 += size2[k] * rot21abs[3 * i + (3 - k - j)] / cl;
    }
    c3 -= rabs(box_dist);
    if (c3 < -margin) return 0;
    if (c3 < separation * ((real)1 - (real)1e-12)) {
      separation = c3; cle1 = 0; cle2 = 0;
      for (int k = 0; k < 3; k++) {
        if (k != i && ((ca[k] > 0) ^ (box_dist < 0))) cle1 += 1 << k;
        if (k != j && ((rot21[3 * i + (3 - k - j)] > 0) ^ (box_dist < 0) ^ ((k - j + 3) % 3 == 1))) cle2 += 1 << k;
      }
      axis_code = 12 + i * 3 + j;
      memcpy(clnorm, ca, sizeof ca);
      inv = box_dist < 0;
    }
  }
  if (axis_code == -1) return 0;
  real points[8][3], depth[8], rotmore[9], rw[9], pw[3], normal[3], hz, rmT[9];
  int n = 0;
  memset(points, 0, sizeof points); memset(depth, 0, sizeof depth);
  if (axis_code < 12) { /* face of one box against the vertices / edges of the other */
    int face_idx = axis_code % 6, box_idx = axis_code / 6;
    rotmore_of(face_idx, rotmore);
    real r[9], p[3], ss[3], rt[9], tmp[3];
    const real* s = box_idx ? size1 : size2;
    mat_mul3(rotmore, box_idx ? rot12 : rot21, r);
    matvec3(rotmore, box_idx ? pos12 : pos21, p);
    matvec3(rotmore, box_idx ? size2 : size1, tmp);
    for (int i = 0; i < 3; i++) ss[i] = rabs(tmp[i]);
    mat_T3(r, rt);
    real lx = ss[0], ly = ss[1];
    hz = ss[2];
    p[2] -= hz;
    int clcorner = 0;
    for (int i = 0; i < 3; i++) if (r[6 + i] < 0) clcorner += 1 << i;
    real lp[3] = {p[0], p[1], p[2]};
    for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) lp[k] += rt[3 * i + k] * s[i] * ((clcorner & (1 << i)) ? (real)1 : (real)-1);
    int dirs = 0;
    real cn1[3] = {0, 0, 0}, cn2[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++) {
      if (rabs(r[6 + i]) < (real)0.5) {
        real* cn = dirs ? cn2 : cn1;
        for (int k = 0; k < 3; k++) cn[k] = rt[3 * i + k] * s[i] * ((clcorner & (1 << i)) ? (real)-2 : (real)2);
        dirs++;
      }
    }
    int kk = dirs * dirs;
    for (int i = 0; i < kk; i++) for (int q = 0; q < 2; q++) { /* incident-face edges against the reference rectangle's sides */
      real lav[3], lbv[3];
      for (int k = 0; k < 3; k++) { lav[k] = lp[k] + (i < 2 ? 0 : (i == 2 ? cn1[k] : cn2[k])); lbv[k] = (i == 0 || i == 3) ? cn1[k] : cn2[k]; }
      if (rabs(lbv[q]) > MJ_MINVAL) {
        real br = 1 / lbv[q];
        for (int j = -1; j <= 1; j += 2) {
          real l = ss[q] * (real)j, c1 = (l - lav[q]) * br;
          if (c1 < 0 || c1 > 1) continue;
          real c2 = lav[1 - q] + lbv[1 - q] * c1;
          if (rabs(c2) > ss[1 - q]) continue;
          if (n < 8) { for (int k = 0; k < 3; k++) points[n][k] = lav[k] + c1 * lbv[k]; n++; }
        }
      }
    }
    if (dirs == 2) { /* reference-face corners inside the incident parallelogram */
      real ax = cn1[0], bx = cn2[0], ay = cn1[1], by = cn2[1], C = safe_div(1, ax * by - bx * ay);
      for (int i = 0; i < 4; i++) {
        real llx = (i / 2) ? lx : -lx, lly = (i % 2) ? ly : -ly, x = llx - lp[0], y = lly - lp[1];
        real u = (x * by - y * bx) * C, v = (y * ax - x * ay) * C;
        if (u > 0 && v > 0 && u < 1 && v < 1 && n < 8) { points[n][0] = llx; points[n][1] = lly; points[n][2] = lp[2] + u * cn1[2] + v * cn2[2]; n++; }
      }
    }
    for (int i = 0; i < (1 << dirs); i++) { /* incident corners inside the reference rectangle */
      real t[3];
      for (int k = 0; k < 3; k++) t[k] = lp[k] + (real)(i & 1) * cn1[k] + (real)((i & 2) != 0) * cn2[k];
      if (t[0] > -lx && t[0] < lx && t[1] > -ly && t[1] < ly && n < 8) { memcpy(points[n], t, sizeof t); n++; }
    }
    int m = n; n = 0;
    for (int i = 0; i < m; i++) {
      if (points[i][2] > margin) continue;
      if (i != n) memcpy(points[n], points[i], 3 * sizeof(real));
      depth[n] = points[n][2];
      points[n][2] *= (real)0.5;
      n++;
    }
    mat_T3(rotmore, rmT);
    mat_mul3(box_idx ? rot2 : rot1, rmT, rw);
    memcpy(pw, box_idx ? pos2 : pos1, 3 * sizeof(real));
    for (int k = 0; k < 3; k++) normal[k] = (box_idx ? (real)-1 : (real)1) * rw[3 * k + 2];
  } else { /* edge of box1 against edge of box2 */
    int edge1 = (axis_code - 12) / 3, edge2 = (axis_code - 12) % 3;
    int ax1 = 1 - (edge2 & 1), ax2 = 2 - (edge2 & 2), pax1 = 1 - (edge1 & 1), pax2 = 2 - (edge1 & 2);
    if (rot21abs[3 * edge1 + ax1] < rot21abs[3 * edge1 + ax2]) { int t = ax1; ax1 = ax2; ax2 = t; }
    if (rot12abs[3 * edge2 + pax1] < rot12abs[3 * edge2 + pax2]) { int t = pax1; pax1 = pax2; pax2 = t; }
    rotmore_of((cle1 & (1 << pax2)) ? pax2 : pax2 + 3, rotmore);
    real p[3], rnorm[3], r[9], rt[9], s[3], tmp[3];
    matvec3(rotmore, pos21, p);
    matvec3(rotmore, clnorm, rnorm);
    mat_mul3(rotmore, rot21, r);
    mat_T3(r, rt);
    mat_T3(rotmore, rmT);
    matvec3(rmT, size1, tmp);
    for (int i = 0; i < 3; i++) s[i] = rabs(tmp[i]);
    real lx = s[0], ly = s[1];
    hz = s[2];
    p[2] -= hz;
    real sg1 = (cle2 & (1 << ax1)) ? (real)1 : (real)-1, sg2 = (cle2 & (1 << ax2)) ? (real)1 : (real)-1;
    for (int k = 0; k < 3; k++) {
      real base0 = p[k] + rt[3 * ax1 + k] * size2[ax1] * sg1 + rt[3 * ax2 + k] * size2[ax2] * sg2;
      real base2 = p[k] - rt[3 * ax1 + k] * size2[ax1] * sg1 + rt[3 * ax2 + k] * size2[ax2] * sg2;
      real e = rt[3 * edge2 + k] * size2[edge2];
      points[0][k] = base0 + e; points[1][k] = base0 - e; points[2][k] = base2 + e; points[3][k] = base2 - e;
    }
    real axi_lp[3], axi_cn1[3], axi_cn2[3];
    for (int k = 0; k < 3; k++) { axi_lp[k] = points[0][k]; axi_cn1[k] = points[1][k] - points[0][k]; axi_cn2[k] = points[2][k] - points[0][k]; }
    if (rabs(rnorm[2]) < MJ_MINVAL) return 0;
    real sgn = inv ? (real)-1 : (real)1, innorm = sgn / rnorm[2], pu[4][3];
    for (int i = 0; i < 4; i++) {
      memcpy(pu[i], points[i], 3 * sizeof(real));
      real c_scl = points[i][2] * sgn * innorm;
      for (int k = 0; k < 3; k++) points[i][k] -= rnorm[k] * c_scl;
    }
    real pts_lp[3], pts_cn1[3], pts_cn2[3];
    for (int k = 0; k < 3; k++) { pts_lp[k] = points[0][k]; pts_cn1[k] = points[1][k] - points[0][k]; pts_cn2[k] = points[2][k] - points[0][k]; }
    n = 0;
    for (int i = 0; i < 4; i++) for (int q = 0; q < 2; q++) {
      real la = pts_lp[q] + (i < 2 ? 0 : (i == 2 ? pts_cn1[q] : pts_cn2[q])), lb = (i == 0 || i == 3) ? pts_cn1[q] : pts_cn2[q];
      real lc = pts_lp[1 - q] + (i < 2 ? 0 : (i == 2 ? pts_cn1[1 - q] : pts_cn2[1 - q])), ld = (i == 0 || i == 3) ? pts_cn1[1 - q] : pts_cn2[1 - q];
      real lua[3], lub[3];
      for (int k = 0; k < 3; k++) { lua[k] = axi_lp[k] + (i < 2 ? 0 : (i == 2 ? axi_cn1[k] : axi_cn2[k])); lub[k] = (i == 0 || i == 3) ? axi_cn1[k] : axi_cn2[k]; }
      if (rabs(lb) > MJ_MINVAL) {
        real br = 1 / lb;
        for (int j = -1; j <= 1; j += 2) {
          if (n == 8) break;
          real l = s[q] * (real)j, c1 = (l - la) * br;
          if (c1 < 0 || c1 > 1) continue;
          real c2 = lc + ld * c1;
          if (rabs(c2) > s[1 - q]) continue;
          if ((lua[2] + lub[2] * c1) * innorm > margin) continue;
          for (int k = 0; k < 3; k++) points[n][k] = lua[k] * (real)0.5 + c1 * lub[k] * (real)0.5;
          points[n][q] += (real)0.5 * l;
          points[n][1 - q] += (real)0.5 * c2;
          depth[n] = points[n][2] * innorm * 2;
          n++;
        }
      }
    }
    int nl = n;
    real ax = pts_cn1[0], bx = pts_cn2[0], ay = pts_cn1[1], by = pts_cn2[1], C = safe_div(1, ax * by - bx * ay);
    for (int i = 0; i < 4; i++) {
      if (n == 8) break;
      real llx = (i / 2) ? lx : -lx, lly = (i % 2) ? ly : -ly, x = llx - pts_lp[0], y = lly - pts_lp[1];
      real u = (x * by - y * bx) * C, v = (y * ax - x * ay) * C;
      if (nl == 0) { if ((u < 0 || u > 1) && (v < 0 || v > 1)) continue; }
      else if (u < 0 || v < 0 || u > 1 || v > 1) continue;
      u = rclamp(u, 0, 1); v = rclamp(v, 0, 1);
      real wgt = 1 - u - v, vtmp[3], pt[3] = {llx, lly, 0}, tc1 = 0;
      for (int k = 0; k < 3; k++) { vtmp[k] = pu[0][k] * wgt + pu[1][k] * u + pu[2][k] * v; tc1 += (pt[k] - vtmp[k]) * (pt[k] - vtmp[k]); }
      if (vtmp[2] > 0 && tc1 > margin * margin) continue;
      for (int k = 0; k < 3; k++) points[n][k] = (real)0.5 * (pt[k] + vtmp[k]);
      depth[n] = (real)sqrt((double)tc1) * (vtmp[2] < 0 ? (real)-1 : (real)1);
      n++;
    }
    int nf = n;
    for (int i = 0; i < 4; i++) {
      if (n >= 8) break;
      real x = pu[i][0], y = pu[i][1];
      if (nl == 0 && nf != 0) { if ((x < -lx || x > lx) && (y < -ly || y > ly)) continue; }
      else if (x < -lx || x > lx || y < -ly || y > ly) continue;
      real c1 = 0;
      for (int j = 0; j < 2; j++) {
        if (pu[i][j] < -s[j]) c1 += (pu[i][j] + s[j]) * (pu[i][j] + s[j]);
        else if (pu[i][j] > s[j]) c1 += (pu[i][j] - s[j]) * (pu[i][j] - s[j]);
      }
      c1 += pu[i][2] * innorm * pu[i][2] * innorm;
      if (pu[i][2] > 0 && c1 > margin * margin) continue;
      real tp[3] = {pu[i][0], pu[i][1], 0};
      for (int j = 0; j < 2; j++) {
        if (pu[i][j] < -s[j]) tp[j] = -s[j] * (real)0.5;
        else if (pu[i][j] > s[j]) tp[j] = s[j] * (real)0.5;
      }
      for (int k = 0; k < 3; k++) points[n][k] = (tp[k] + pu[i][k]) * (real)0.5;
      depth[n] = (real)sqrt((double)c1) * (pu[i][2] < 0 ? (real)-1 : (real)1);
      n++;
    }
    mat_mul3(rot1, rmT, rw);
    memcpy(pw, pos1, 3 * sizeof(real));
    real nn[3]; matvec3(rw, rnorm, nn);
    for (int k = 0; k < 3; k++) normal[k] = sgn * nn[k];
  }
  for (int i = 0; i < n; i++) {
    points[i][2] += hz;
    real g[3]; matvec3(rw, points[i], g);
    for (int k = 0; k < 3; k++) cpos[i][k] = g[k] + pw[k];
    cdist[i] = depth[i];
  }
  memcpy(cnormal, normal, sizeof normal);
  return n;
}
#include "oracle_ccd.h"

/* collision_driver.py:47-81: pair types the reference sends to the convex (GJK / EPA) path, in MJ_COLLISION_TABLE order,
 * restricted to analytic geoms.  Box-box is convex unless the nativeccd disable flag routes it to the primitive. */
static const int CONVEX_PAIRS[][2] = {{GEOM_SPHERE, GEOM_ELLIPSOID}, {GEOM_SPHERE, GEOM_MESH}, {GEOM_CAPSULE, GEOM_ELLIPSOID}, {GEOM_CAPSULE, GEOM_CYLINDER},
  {GEOM_CAPSULE, GEOM_MESH}, {GEOM_ELLIPSOID, GEOM_ELLIPSOID}, {GEOM_ELLIPSOID, GEOM_CYLINDER}, {GEOM_ELLIPSOID, GEOM_BOX}, {GEOM_ELLIPSOID, GEOM_MESH},
  {GEOM_CYLINDER, GEOM_CYLINDER}, {GEOM_CYLINDER, GEOM_BOX}, {GEOM_CYLINDER, GEOM_MESH}, {GEOM_BOX, GEOM_BOX}, {GEOM_BOX, GEOM_MESH}, {GEOM_MESH, GEOM_MESH}};
#define N_CONVEX_PAIRS ((int)(sizeof CONVEX_PAIRS / sizeof CONVEX_PAIRS[0]))
/* box-box is a convex pair unless the model's nativeccd disable flag is set */
static int convex_pair_rank(int t1, int t2, int nativeccd) {
  if (t1 == GEOM_BOX && t2 == GEOM_BOX && !nativeccd) return -1;
  for (int i = 0; i < N_CONVEX_PAIRS; i++) if (CONVEX_PAIRS[i][0] == t1 && CONVEX_PAIRS[i][1] == t2) return i;
  return -1;
}
/* collision_core.py:60-140 geom(): pose, size and -- for meshes -- the asset's vertex block, hull graph and polygon tables */
static void fill_cgeom(const W* w, int g, real margin, CGeom* c) {
  const OrcModel* m = w->m;
  memset(c, 0, sizeof *c);
  memcpy(c->pos, w->geom_xpos + 3 * g, sizeof c->pos); memcpy(c->rot, w->geom_xmat + 9 * g, sizeof c->rot); memcpy(c->size, m->geom_size + 3 * g, sizeof c->size);
  c->type = m->geom_type[g]; c->margin = margin; c->index = -1;
  if (c->type == GEOM_MESH && m->geom_dataid[g] >= 0) {
    const int id = m->geom_dataid[g], vadr = m->mesh_vertadr[id], padr = m->mesh_polyadr[id];
    c->vert = m->mesh_vert + 3 * vadr; c->vertnum = m->mesh_vertnum[id];
    c->graph = m->mesh_graphadr[id] >= 0 ? m->mesh_graph + m->mesh_graphadr[id] : NULL;
    c->polynum = m->mesh_polynum[id]; c->polynormal = m->mesh_polynormal + 3 * padr;
    c->polyvertadr = m->mesh_polyvertadr + padr; c->polyvertnum = m->mesh_polyvertnum + padr; c->polyvert = m->mesh_polyvert;
    c->polymapadr = m->mesh_polymapadr + vadr; c->polymapnum = m->mesh_polymapnum + vadr; c->polymap = m->mesh_polymap;
  }
}
/* collision_convex.py:739-968 eval_ccd_write_contact (single contact; multi-contact applies to box / mesh pairs only) */
static void convex_pair(W* w, int g1, int g2, int pairid) {
  const OrcModel* m = w->m;
  ConParams p;
  contact_params(m, g1, g2, pairid, &p);
  CGeom a, b;
  fill_cgeom(w, g1, p.margin, &a); fill_cgeom(w, g2, p.margin, &b);
  real dist, w1[4][3], w2[4][3], frame[9], nrm[3], pos[3]; int ovf = 0;
  memset(w1, 0, sizeof w1); memset(w2, 0, sizeof w2);
  int ncon = ccd_pair(m->ccd_tolerance, p.gap, m->ccd_iterations, m->epa_iterations, 1, a, b, &dist, w1, w2, &ovf);
  if (ovf) w->overflow[0] |= OVF_EPA_HORIZON;
  if (ncon == 0 || dist >= p.gap) return;
  dist += p.margin; /* back to the distance between the un-inflated surfaces (collision_convex.py:862-868) */
  if (dist <= p.margin) v3sub(w1[0], w2[0], nrm); else v3sub(w2[0], w1[0], nrm);
  make_frame(nrm, frame);
  for (int k = 0; k < ncon; k++) {
    for (int i = 0; i < 3; i++) pos[i] = (real)0.5 * (w1[k][i] + w2[k][i]);
    write_contact(w, k, dist, pos, frame, &p, g1, g2);
  }
}

/* Direct entry to the convex pair routine, shaped like the harness of the reference's own GJK tests
 * (collision_gjk_test.py:35-303 _geom_dist: ccd() with the same iteration count for GJK and EPA, optional multicontact). */
int orc_ccd(int type1, const real* size1, const real* pos1, const real* mat1, int type2, const real* size2, const real* pos2, const real* mat2,
            real margin, real tolerance, real cutoff, int iterations, int multi, real* dist, real* w1, real* w2, int* overflow) {
  CGeom a, b;
  memset(&a, 0, sizeof a); memset(&b, 0, sizeof b); a.index = b.index = -1;
  memcpy(a.pos, pos1, sizeof a.pos); memcpy(a.rot, mat1, sizeof a.rot); memcpy(a.size, size1, sizeof a.size); a.type = type1; a.margin = margin;
  memcpy(b.pos, pos2, sizeof b.pos); memcpy(b.rot, mat2, sizeof b.rot); memcpy(b.size, size2, sizeof b.size); b.type = type2; b.margin = margin;
  real x1[4][3], x2[4][3]; int ovf = 0;
  memset(x1, 0, sizeof x1); memset(x2, 0, sizeof x2);
  int n = ccd_pair(tolerance, cutoff, iterations, iterations, multi, a, b, dist, x1, x2, &ovf);
  memcpy(w1, x1, sizeof x1); memcpy(w2, x2, sizeof x2);
  *overflow = ovf;
  return n;
}

/* Same, with the geoms given by descriptor so that mesh geoms (vertex block, hull graph, polygon tables -- already offset to the mesh the way
 * fill_cgeom does it) can be passed; separate GJK / EPA iteration counts. */
typedef struct {
  int type, vertnum, polynum, pad;
  const real *size, *pos, *mat, *vert, *polynormal;
  const int *graph, *polyvertadr, *polyvertnum, *polyvert, *polymapadr, *polymapnum, *polymap;
} OrcGeomDesc;
static void desc_cgeom(const OrcGeomDesc* d, real margin, CGeom* c) {
  memset(c, 0, sizeof *c);
  memcpy(c->pos, d->pos, sizeof c->pos); memcpy(c->rot, d->mat, sizeof c->rot); memcpy(c->size, d->size, sizeof c->size);
  c->type = d->type; c->margin = margin; c->index = -1;
  c->vertnum = d->vertnum; c->polynum = d->polynum; c->vert = d->vert; c->polynormal = d->polynormal; c->graph = d->graph;
  c->polyvertadr = d->polyvertadr; c->polyvertnum = d->polyvertnum; c->polyvert = d->polyvert;
  c->polymapadr = d->polymapadr; c->polymapnum = d->polymapnum; c->polymap = d->polymap;
}
int orc_ccd_desc(const OrcGeomDesc* d1, const OrcGeomDesc* d2, real margin, real tolerance, real cutoff, int gjk_iterations, int epa_iterations, int multi,
                 real* dist, real* w1, real* w2, int* overflow) {
  CGeom a, b;
  desc_cgeom(d1, margin, &a); desc_cgeom(d2, margin, &b);
  real x1[4][3], x2[4][3]; int ovf = 0;
  memset(x1, 0, sizeof x1); memset(x2, 0, sizeof x2);
  int n = ccd_pair(tolerance, cutoff, gjk_iterations, epa_iterations, multi, a, b, dist, x1, x2, &ovf);
  memcpy(w1, x1, sizeof x1); memcpy(w2, x2, sizeof x2);
  *overflow = ovf;
  return n;
}

static real pc_support(const real* ppl, const real* v, const real* n) { real d[3]; v3sub(ppl, v, d); return dot3(d, n); }
/* collision_primitive.py:52-277 plane_convex: up to four well-spread vertices of the convex geom that lie (nearly) deepest below the plane */
static void plane_convex(const real* n_world, const real* plane_pos, const CGeom* c, real dist[4], real pos[4][3]) {
  const real HUGE_V = (real)1e6;
  real d[3], ppl[3], n[3];
  int idx[4] = {-1, -1, -1, -1};
  for (int i = 0; i < 4; i++) { dist[i] = MJ_MAXVAL; pos[i][0] = pos[i][1] = pos[i][2] = 0; }
  v3sub(plane_pos, c->pos, d); matT_vec3(c->rot, d, ppl); matT_vec3(c->rot, n_world, n);
#define PC_SUPPORT(v) pc_support(ppl, (v), n) /* dot(plane_pos_local - vert, n), evaluated in the reference's order */
  if (!c->graph || c->vertnum < 10) {
    real max_support = -HUGE_V; const real* a = NULL;
    for (int i = 0; i < c->vertnum; i++) { real s = PC_SUPPORT(c->vert + 3 * i); if (s > max_support) { max_support = s; idx[0] = i; a = c->vert + 3 * i; } }
    if (max_support < 0) return;
    real threshold = max_support - (real)1e-3, best = -HUGE_V; const real* b = NULL;
    for (int i = 0; i < c->vertnum; i++) {
      const real* v = c->vert + 3 * i; real df[3]; v3sub(a, v, df);
      real dd = dot3(df, df) + (PC_SUPPORT(v) > threshold ? 0 : -HUGE_V);
      if (dd > best) { idx[1] = i; best = dd; b = v; }
    }
    real amb[3], ab[3]; v3sub(a, b, amb); cross3(n, amb, ab);
    best = -HUGE_V; const real* cc = NULL;
    for (int i = 0; i < c->vertnum; i++) {
      const real* v = c->vert + 3 * i; real ap[3]; v3sub(a, v, ap);
      real dd = rabs(dot3(ap, ab)) + (PC_SUPPORT(v) > threshold ? 0 : -HUGE_V);
      if (dd > best) { idx[2] = i; best = dd; cc = v; }
    }
    real amc[3], bmc[3], ac[3], bc[3]; v3sub(a, cc, amc); v3sub(b, cc, bmc); cross3(n, amc, ac); cross3(n, bmc, bc);
    best = -HUGE_V;
    for (int i = 0; i < c->vertnum; i++) {
      const real* v = c->vert + 3 * i; real ap[3], bp[3]; v3sub(a, v, ap); v3sub(b, v, bp);
      real mask = PC_SUPPORT(v) > threshold ? 0 : -HUGE_V;
      real dd = (rabs(dot3(ap, ac)) + mask) + (rabs(dot3(bp, bc)) + mask);
      if (dd > best) { idx[3] = i; best = dd; }
    }
  } else {
    const int numvert = c->graph[0], *vert_edgeadr = c->graph + 2, *vert_globalid = c->graph + 2 + numvert, *edge_localid = c->graph + 2 + 2 * numvert;
    real max_support = -HUGE_V; int prev, imax = 0;
    for (;;) { /* hill climb to the deepest vertex */
      prev = imax;
      for (int i = vert_edgeadr[imax]; edge_localid[i] >= 0; i++) { int sub = edge_localid[i]; real s = PC_SUPPORT(c->vert + 3 * vert_globalid[sub]); if (s > max_support) { max_support = s; imax = sub; } }
      if (imax == prev) break;
    }
    real threshold = rmax(0, max_support - (real)1e-3), best = -HUGE_V;
    for (;;) {
      prev = imax;
      for (int i = vert_edgeadr[imax]; edge_localid[i] >= 0; i++) { int sub = edge_localid[i]; real s = PC_SUPPORT(c->vert + 3 * vert_globalid[sub]); real dd = s > threshold ? s : -HUGE_V; if (dd > best) { best = dd; imax = sub; } }
      if (imax == prev) break;
    }
    const real* a = c->vert + 3 * vert_globalid[imax]; idx[0] = vert_globalid[imax];
    best = -HUGE_V;
    for (;;) {
      prev = imax;
      for (int i = vert_edgeadr[imax]; edge_localid[i] >= 0; i++) {
        int sub = edge_localid[i]; const real* v = c->vert + 3 * vert_globalid[sub]; real df[3]; v3sub(a, v, df);
        real dd = dot3(df, df) + (PC_SUPPORT(v) > threshold ? 0 : -HUGE_V);
        if (dd > best) { best = dd; imax = sub; }
      }
      if (imax == prev) break;
    }
    const real* b = c->vert + 3 * vert_globalid[imax]; idx[1] = vert_globalid[imax];
    real amb[3], ab[3]; v3sub(a, b, amb); cross3(n, amb, ab);
    best = -HUGE_V;
    for (;;) {
      prev = imax;
      for (int i = vert_edgeadr[imax]; edge_localid[i] >= 0; i++) {
        int sub = edge_localid[i]; const real* v = c->vert + 3 * vert_globalid[sub]; real ap[3]; v3sub(a, v, ap);
        real dd = rabs(dot3(ap, ab)) + (PC_SUPPORT(v) > threshold ? 0 : -HUGE_V);
        if (dd > best) { best = dd; imax = sub; }
      }
      if (imax == prev) break;
    }
    const real* cc = c->vert + 3 * vert_globalid[imax]; idx[2] = vert_globalid[imax];
    real amc[3], bmc[3], ac[3], bc[3]; v3sub(a, cc, amc); v3sub(b, cc, bmc); cross3(n, amc, ac); cross3(n, bmc, bc);
    best = -HUGE_V;
    for (;;) {
      prev = imax;
      for (int i = vert_edgeadr[imax]; edge_localid[i] >= 0; i++) {
        int sub = edge_localid[i]; const real* v = c->vert + 3 * vert_globalid[sub]; real ap[3], bp[3]; v3sub(a, v, ap); v3sub(b, v, bp);
        real mask = PC_SUPPORT(v) > threshold ? 0 : -HUGE_V;
        real dd = (rabs(dot3(ap, ac)) + mask) + (rabs(dot3(bp, bc)) + mask);
        if (dd > best) { best = dd; imax = sub; }
      }
      if (imax == prev) break;
    }
    idx[3] = vert_globalid[imax];
  }
  int count = 0;
  for (int i = 3; i >= 0; i--) { /* unique indices, last first */
    int uniq = 0;
    for (int j = 0; j <= i; j++) if (idx[j] == idx[i]) uniq++;
    if (uniq != 1) continue;
    const real* v = c->vert + 3 * idx[i];
    real wp[3]; matvec3(c->rot, v, wp);
    real dd = -PC_SUPPORT(v);
    for (int k = 0; k < 3; k++) pos[count][k] = c->pos[k] + wp[k] - (real)0.5 * dd * n_world[k];
    dist[count] = dd;
    count++;
  }
#undef PC_SUPPORT
}

/* plane_convex on a described geom (tests of the device routine on the host) */
void orc_plane_convex_desc(const real* n_world, const real* plane_pos, const OrcGeomDesc* d, real* dist, real* pos) {
  CGeom c; desc_cgeom(d, 0, &c);
  real d4[4], p4[4][3];
  plane_convex(n_world, plane_pos, &c, d4, p4);
  memcpy(dist, d4, sizeof d4); memcpy(pos, p4, sizeof p4);
}

static void narrowphase_pair(W* w, int g1, int g2, int pairid) {
  const OrcModel* m = w->m;
  int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
  ConParams p;
  contact_params(m, g1, g2, pairid, &p);
  const real *pos1 = w->geom_xpos + 3 * g1, *pos2 = w->geom_xpos + 3 * g2, *rot1 = w->geom_xmat + 9 * g1, *rot2 = w->geom_xmat + 9 * g2;
  const real *size1 = m->geom_size + 3 * g1, *size2 = m->geom_size + 3 * g2;
  real ax1[3] = {rot1[2], rot1[5], rot1[8]}, ax2[3] = {rot2[2], rot2[5], rot2[8]};
  real frame[9], pos[3], n[3];
  if (t1 == GEOM_PLANE && t2 == GEOM_SPHERE) { /* collision_primitive.py:281 */
    real dist = plane_sphere(ax1, pos1, pos2, size2[0], pos);
    make_frame(ax1, frame);
    write_contact(w, 0, dist, pos, frame, &p, g1, g2);
  } else if (t1 == GEOM_PLANE && t2 == GEOM_CAPSULE) { /* collision_primitive_core.py:252, collision_primitive.py:600 */
    real b[3], c[3], an = dot3(ax1, ax2);
    for (int i = 0; i < 3; i++) b[i] = ax2[i] - ax1[i] * an;
    real bn = len3(b);
    if (bn != 0) for (int i = 0; i < 3; i++) b[i] /= bn;
    if (bn < 0.5) {
      if (-0.5 < ax1[1] && ax1[1] < 0.5) { b[0] = 0; b[1] = 1; b[2] = 0; } else { b[0] = 0; b[1] = 0; b[2] = 1; }
    }
    cross3(ax1, b, c);
    for (int i = 0; i < 3; i++) { frame[i] = ax1[i]; frame[3 + i] = b[i]; frame[6 + i] = c[i]; }
    real e1[3], e2[3], pp1[3], pp2[3];
    for (int i = 0; i < 3; i++) { e1[i] = pos2[i] + ax2[i] * size2[1]; e2[i] = pos2[i] - ax2[i] * size2[1]; }
    real d1 = plane_sphere(ax1, pos1, e1, size2[0], pp1), d2 = plane_sphere(ax1, pos1, e2, size2[0], pp2);
    write_contact(w, 0, d1, pp1, frame, &p, g1, g2);
    write_contact(w, 1, d2, pp2, frame, &p, g1, g2);
  } else if (t1 == GEOM_SPHERE && t2 == GEOM_SPHERE) {
    real dist = sphere_sphere(pos1, size1[0], pos2, size2[0], pos, n);
    make_frame(n, frame);
    write_contact(w, 0, dist, pos, frame, &p, g1, g2);
  } else if (t1 == GEOM_SPHERE && t2 == GEOM_CAPSULE) { /* collision_primitive_core.py:87 */
    real a[3], b[3], pt[3];
    for (int i = 0; i < 3; i++) { a[i] = pos2[i] - ax2[i] * size2[1]; b[i] = pos2[i] + ax2[i] * size2[1]; }
    orc_closest_segment_point(a, b, pos1, pt);
    real dist = sphere_sphere(pos1, size1[0], pt, size2[0], pos, n);
    make_frame(n, frame);
    write_contact(w, 0, dist, pos, frame, &p, g1, g2);
  } else if (t1 == GEOM_CAPSULE && t2 == GEOM_CAPSULE) { /* collision_primitive_core.py:122-249 */
    real axis1[3], axis2[3], dif[3], margin = p.margin;
    for (int i = 0; i < 3; i++) { axis1[i] = ax1[i] * size1[1]; axis2[i] = ax2[i] * size2[1]; dif[i] = pos1[i] - pos2[i]; }
    real ma = dot3(axis1, axis1), mb = -dot3(axis1, axis2), mc = dot3(axis2, axis2), u = -dot3(axis1, dif), v = dot3(axis2, dif);
    real det = ma * mc - mb * mb, v1[3], v2[3];
    real cdist[2] = {INFINITY, INFINITY}, cpos[2][3] = {{0}}, cn[2][3] = {{0}};
    if (fabs((double)det) >= MJ_MINVAL) {
      real inv = 1 / det, x1 = (mc * u - mb * v) * inv, x2 = (ma * v - mb * u) * inv;
      if (x1 > 1) { x1 = 1; x2 = (v - mb) / mc; } else if (x1 < -1) { x1 = -1; x2 = (v + mb) / mc; }
      if (x2 > 1) { x2 = 1; x1 = rclamp((u - mb) / ma, -1, 1); } else if (x2 < -1) { x2 = -1; x1 = rclamp((u + mb) / ma, -1, 1); }
      for (int i = 0; i < 3; i++) { v1[i] = pos1[i] + axis1[i] * x1; v2[i] = pos2[i] + axis2[i] * x2; }
      real dist = sphere_sphere(v1, size1[0], v2, size2[0], pos, n);
      if (dist <= margin) { cdist[0] = dist; memcpy(cpos[0], pos, sizeof pos); memcpy(cn[0], n, sizeof n); }
    } else {
      int cc = 0; real x1, x2, dist;
      for (int i = 0; i < 3; i++) v1[i] = pos1[i] + axis1[i];
      x2 = rclamp((v - mb) / mc, -1, 1);
      for (int i = 0; i < 3; i++) v2[i] = pos2[i] + axis2[i] * x2;
      dist = sphere_sphere(v1, size1[0], v2, size2[0], pos, n);
      if (dist <= margin) { cdist[cc] = dist; memcpy(cpos[cc], pos, sizeof pos); memcpy(cn[cc], n, sizeof n); cc++; }
      for (int i = 0; i < 3; i++) v1[i] = pos1[i] - axis1[i];
      x2 = rclamp((v + mb) / mc, -1, 1);
      for (int i = 0; i < 3; i++) v2[i] = pos2[i] + axis2[i] * x2;
      dist = sphere_sphere(v1, size1[0], v2, size2[0], pos, n);
      if (dist <= margin) { cdist[cc] = dist; memcpy(cpos[cc], pos, sizeof pos); memcpy(cn[cc], n, sizeof n); cc++; }
      if (cc < 2) {
        for (int i = 0; i < 3; i++) v2[i] = pos2[i] + axis2[i];
        x1 = rclamp((u - mb) / ma, -1, 1);
        for (int i = 0; i < 3; i++) v1[i] = pos1[i] + axis1[i] * x1;
        dist = sphere_sphere(v1, size1[0], v2, size2[0], pos, n);
        if (dist <= margin) { cdist[cc] = dist; memcpy(cpos[cc], pos, sizeof pos); memcpy(cn[cc], n, sizeof n); cc++; }
      }
      if (cc < 2) {
        for (int i = 0; i < 3; i++) v2[i] = pos2[i] - axis2[i];
        x1 = rclamp((u + mb) / ma, -1, 1);
        for (int i = 0; i < 3; i++) v1[i] = pos1[i] + axis1[i] * x1;
        dist = sphere_sphere(v1, size1[0], v2, size2[0], pos, n);
        if (dist <= margin) { cdist[cc] = dist; memcpy(cpos[cc], pos, sizeof pos); memcpy(cn[cc], n, sizeof n); }
      }
    }
    for (int i = 0; i < 2; i++) { make_frame(cn[i], frame); write_contact(w, i, cdist[i], cpos[i], frame, &p, g1, g2); }
  } else if (t1 == GEOM_PLANE && t2 == GEOM_ELLIPSOID) { /* collision_primitive.py:686 */
    real dist = plane_ellipsoid(ax1, pos1, pos2, rot2, size2, pos);
    make_frame(ax1, frame);
    write_contact(w, 0, dist, pos, frame, &p, g1, g2);
  } else if (t1 == GEOM_PLANE && t2 == GEOM_CYLINDER) { /* collision_primitive.py:1000 */
    real d4[4], p4[4][3];
    plane_cylinder(ax1, pos1, pos2, ax2, size2[0], size2[1], d4, p4);
    make_frame(ax1, frame);
    for (int i = 0; i < 4; i++) write_contact(w, i, d4[i], p4[i], frame, &p, g1, g2);
  } else if (t1 == GEOM_PLANE && t2 == GEOM_MESH) { /* collision_primitive.py:838 plane_convex_wrapper */
    CGeom c; fill_cgeom(w, g2, 0, &c);
    real d4[4], p4[4][3];
    plane_convex(ax1, pos1, &c, d4, p4);
    make_frame(ax1, frame);
    for (int i = 0; i < 4; i++) write_contact(w, i, d4[i], p4[i], frame, &p, g1, g2);
  } else if (t1 == GEOM_PLANE && t2 == GEOM_BOX) { /* collision_primitive.py:761 */
    real d8[8], p8[8][3];
    plane_box(ax1, pos1, pos2, rot2, size2, d8, p8);
    make_frame(ax1, frame);
    for (int i = 0; i < 8; i++) write_contact(w, i, d8[i], p8[i], frame, &p, g1, g2);
  } else if (t1 == GEOM_SPHERE && t2 == GEOM_CYLINDER) { /* collision_primitive.py:915 */
    real dist = sphere_cylinder(pos1, size1[0], pos2, ax2, size2[0], size2[1], pos, n);
    make_frame(n, frame);
    write_contact(w, 0, dist, pos, frame, &p, g1, g2);
  } else if (t1 == GEOM_SPHERE && t2 == GEOM_BOX) { /* collision_primitive.py:1087 */
    real dist = sphere_box(pos1, size1[0], pos2, rot2, size2, pos, n);
    make_frame(n, frame);
    write_contact(w, 0, dist, pos, frame, &p, g1, g2);
  } else if (t1 == GEOM_CAPSULE && t2 == GEOM_BOX) { /* collision_primitive.py:1161 */
    real d2[2], p2[2][3], n2[2][3];
    capsule_box(pos1, ax1, size1[0], size1[1], pos2, rot2, size2, d2, p2, n2);
    for (int i = 0; i < 2; i++) { make_frame(n2[i], frame); write_contact(w, i, d2[i], p2[i], frame, &p, g1, g2); }
  } else if (t1 == GEOM_BOX && t2 == GEOM_BOX && (m->disableflags & DSBL_NATIVECCD)) { /* collision_primitive.py:1250; primitive only with nativeccd off (collision_driver.py:868) */
    real d8[8], p8[8][3], nn[3];
    int nc = box_box(pos1, rot1, size1, pos2, rot2, size2, p.margin, d8, p8, nn);
    make_frame(nn, frame);
    for (int i = 0; i < nc; i++) write_contact(w, i, d8[i], p8[i], frame, &p, g1, g2);
  } else {
    w->overflow[0] |= OVF_UNSUPPORTED;
  }
}
/* collision_driver.py:582-682 sap_broadphase: candidate pairs from a sort of the bounding-sphere projections on a fixed axis
 * (:602-603), in the order the reference's sweep kernel visits them when its threads run one after the other: by world,
 * sorted position of the first geom, sorted position of the second.  Returns the candidate count; cand holds indices into
 * the filtered pair list. */
typedef struct { real lower; int geom; } SapKey;
static int sap_cmp(const void* a, const void* b) {
  const SapKey *x = (const SapKey*)a, *y = (const SapKey*)b;
  if (x->lower < y->lower) return -1;
  if (x->lower > y->lower) return 1;
  return x->geom - y->geom; /* stable */
}
static int sap_candidates(const W* w, int* cand) {
  const OrcModel* m = w->m;
  const int ng = m->ngeom;
  real dir[3] = {(real)0.5935, (real)0.7790, (real)0.1235};
  normalize3(dir);
  SapKey* keys = (SapKey*)malloc((size_t)(ng > 0 ? ng : 1) * sizeof(SapKey));
  real* upper = (real*)malloc((size_t)(ng > 0 ? ng : 1) * sizeof(real));
  int* lookup = (int*)malloc((size_t)(ng > 0 ? ng * ng : 1) * sizeof(int));
  for (int k = 0; k < ng * ng; k++) lookup[k] = -1;
  for (int e = 0; e < m->nxn_npair; e++) lookup[m->nxn_geom_pair[2 * e] * ng + m->nxn_geom_pair[2 * e + 1]] = e;
  for (int g = 0; g < ng; g++) { /* :373-413 sap_project */
    real rbound = m->geom_rbound[g];
    if (rbound == 0) rbound = MJ_MAXVAL;
    real radius = rbound + m->geom_margin[g] + m->geom_gap[g], center = dot3(dir, w->geom_xpos + 3 * g);
    keys[g].geom = g;
    if (center == center) { keys[g].lower = center - radius; upper[g] = center + radius; } else { keys[g].lower = MJ_MAXVAL; upper[g] = MJ_MAXVAL; }
  }
  qsort(keys, (size_t)ng, sizeof(SapKey), sap_cmp);
  int n = 0;
  for (int si = 0; si < ng; si++) {
    /* collision_core.py:489-519: first sorted position past si whose lower bound exceeds this geom's upper bound, clamped to ng - 1
     * (the sweep therefore also visits that first non-overlapping neighbour; the pair filters discard it) */
    int lo = si + 1, hi = ng;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (keys[mid].lower > upper[keys[si].geom]) hi = mid; else lo = mid + 1; }
    int limit = hi < ng - 1 ? hi : ng - 1;
    for (int sj = si + 1; sj <= limit; sj++) {
      int g1 = keys[si].geom, g2 = keys[sj].geom;
      if (g2 < g1) { int t = g1; g1 = g2; g2 = t; }
      int e = lookup[g1 * ng + g2];
      if (e >= 0) cand[n++] = e;
    }
  }
  free(keys); free(upper); free(lookup);
  return n;
}
/* collision_driver.py:884-942 with the NXN (:684-770) or SAP broadphase; contacts are written in candidate order */
static void collision(W* w) {
  const OrcModel* m = w->m;
  w->ncon[0] = 0; w->ncollision[0] = 0;
  if (w->nconmax == 0 || (m->disableflags & (DSBL_CONSTRAINT | DSBL_CONTACT))) return;
  /* Contact order of the reference under sequential execution: the convex narrowphase runs first, one launch per pair type
   * in table order (collision_driver.py:877, collision_convex.py:1369), then the primitive narrowphase; within a launch,
   * broadphase output order. */
  const int nativeccd = !(m->disableflags & DSBL_NATIVECCD);
  int* cand = (int*)malloc((size_t)(m->nxn_npair > 0 ? m->nxn_npair : 1) * sizeof(int));
  int ncand = 0, npass = 0;
  if (m->broadphase == 0) for (int e = 0; e < m->nxn_npair; e++) cand[ncand++] = e;
  else ncand = sap_candidates(w, cand);
  for (int k = 0; k < ncand; k++) {
    int e = cand[k], g1 = m->nxn_geom_pair[2 * e], g2 = m->nxn_geom_pair[2 * e + 1];
    if (!(broadphase_filter(w, g1, g2) || m->nxn_pairid[2 * e + 1] >= 0)) continue;
    w->ncollision[0]++;
    if (m->nxn_pairid[2 * e] == -2) continue; /* sensor-only pair: no constraint contact */
    cand[npass++] = e;
  }
  for (int rank = 0; rank <= N_CONVEX_PAIRS; rank++) {
    for (int k = 0; k < npass; k++) {
      int e = cand[k], g1 = m->nxn_geom_pair[2 * e], g2 = m->nxn_geom_pair[2 * e + 1];
      if (m->geom_type[g1] > m->geom_type[g2]) { int t = g1; g1 = g2; g2 = t; }
      int cr = convex_pair_rank(m->geom_type[g1], m->geom_type[g2], nativeccd);
      if (rank < N_CONVEX_PAIRS) { if (cr == rank) convex_pair(w, g1, g2, m->nxn_pairid[2 * e]); }
      else if (cr < 0) narrowphase_pair(w, g1, g2, m->nxn_pairid[2 * e]);
    }
  }
  free(cand);
  if (w->ncon[0] > w->nconmax) w->ncon[0] = w->nconmax;
}

/* ------------------------------------------------------------------ make_constraint (constraint.py) */
/* constraint.py:83-152 */
static void efc_row(W* w, int efcid, real pos_aref, real pos_imp, real invweight, const real* solref, const real* solimp, real margin, real vel, real frictionloss, int type, int id) {
  const OrcModel* m = w->m;
  real timeconst = solref[0], dampratio = solref[1], dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  if (!(m->disableflags & DSBL_REFSAFE)) timeconst = rmax(timeconst, 2 * m->timestep);
  dmin = rclamp(dmin, MJ_MINIMP, MJ_MAXIMP); dmax = rclamp(dmax, MJ_MINIMP, MJ_MAXIMP);
  width = rmax(MJ_MINVAL, width); mid = rclamp(mid, MJ_MINIMP, MJ_MAXIMP); power = rmax(1, power);
  real dmax_sq = dmax * dmax;
  real k = 1 / (dmax_sq * timeconst * timeconst * dampratio * dampratio), b = 2 / (dmax * timeconst);
  if (solref[0] <= 0) k = -solref[0] / dmax_sq;
  if (solref[1] <= 0) b = -solref[1] / dmax;
  real imp_x = (real)fabs((double)pos_imp) / width;
  real imp_a = (1 / (real)pow((double)mid, (double)power - 1)) * (real)pow((double)imp_x, (double)power);
  real imp_b = 1 - (1 / (real)pow(1 - (double)mid, (double)power - 1)) * (real)pow(1 - (double)imp_x, (double)power);
  real imp_y = imp_x < mid ? imp_a : imp_b;
  real imp = dmin + imp_y * (dmax - dmin);
  imp = rclamp(imp, dmin, dmax);
  if (imp_x > 1) imp = dmax;
  w->efc_D[efcid] = 1 / rmax(invweight * (1 - imp) / imp, MJ_MINVAL);
  w->efc_vel[efcid] = vel;
  w->efc_aref[efcid] = -k * imp * pos_aref - b * vel;
  w->efc_pos[efcid] = pos_aref + margin;
  w->efc_margin[efcid] = margin;
  w->efc_frictionloss[efcid] = frictionloss;
  w->efc_type[efcid] = type;
  w->efc_id[efcid] = id;
}
/* support.py:506 jac_dof / :615 jac_dot_dof: translational and rotational Jacobian column (and its time derivative)
 * of a point attached to body `b` for dof `d`; cvel / cdof_dot are whatever the last velocity stage left in Data */
static void jac_dof(const W* w, const real* point, int b, int d, real* jp, real* jr) {
  const OrcModel* m = w->m;
  for (int i = 0; i < 3; i++) { jp[i] = 0; jr[i] = 0; }
  if (!m->body_isdofancestor[b * m->nv + d]) return;
  const real* com = w->subtree_com + 3 * m->body_rootid[b]; const real* cd = w->cdof + 6 * d;
  real off[3] = {point[0] - com[0], point[1] - com[1], point[2] - com[2]}, cr[3];
  cross3(cd, off, cr);
  for (int i = 0; i < 3; i++) { jp[i] = cd[3 + i] + cr[i]; jr[i] = cd[i]; }
}
static void jac_dot_dof(const W* w, const real* point, int b, int d, real* jp, real* jr) {
  const OrcModel* m = w->m;
  for (int i = 0; i < 3; i++) { jp[i] = 0; jr[i] = 0; }
  if (!m->body_isdofancestor[b * m->nv + d]) return;
  const real* com = w->subtree_com + 3 * m->body_rootid[b];
  real off[3] = {point[0] - com[0], point[1] - com[1], point[2] - com[2]};
  const real* cvel = w->cvel + 6 * b; const real* cd = w->cdof + 6 * d;
  real t[3], pvel[3], cdd[6];
  cross3(off, cvel, t);
  for (int i = 0; i < 3; i++) pvel[i] = cvel[3 + i] - t[i];
  memcpy(cdd, w->cdof_dot + 6 * d, sizeof cdd);
  int j = m->dof_jntid[d], jt = m->jnt_type[j];
  if (jt == JNT_BALL || (jt == JNT_FREE && d >= m->jnt_dofadr[j] + 3)) motion_cross(w->cvel + 6 * m->dof_bodyid[d], cd, cdd);
  real c1[3], c2[3];
  cross3(cdd, off, c1);
  cross3(cd, pvel, c2);
  for (int i = 0; i < 3; i++) { jp[i] = cdd[3 + i] + c1[i] + c2[i]; jr[i] = cdd[i]; }
}
static inline void quat_mul_axis(const real* q, const real* a, real* o) { /* math.py:33 */
  o[0] = -q[1] * a[0] - q[2] * a[1] - q[3] * a[2];
  o[1] = q[0] * a[0] + q[2] * a[2] - q[3] * a[1];
  o[2] = q[0] * a[1] + q[3] * a[0] - q[1] * a[2];
  o[3] = q[0] * a[2] + q[1] * a[1] - q[2] * a[0];
}
/* constraint.py:156 connect (3 rows), :966 weld (6 rows), :500 joint (1 row); body-based anchors only */
static int equality_rows(W* w, int nefc) {
  const OrcModel* m = w->m; const int nv = m->nv, njmax = w->njmax;
  for (int pass = 0; pass < 4; pass++) for (int e = 0; e < m->neq; e++) { /* launch order: connect, weld, joint, tendon */
    int type = m->eq_type[e];
    if (type != (pass == 0 ? EQ_CONNECT : pass == 1 ? EQ_WELD : pass == 2 ? EQ_JOINT : EQ_TENDON) || !w->eq_active[e]) continue;
    const real* data = m->eq_data + 11 * e; const real* solref = m->eq_solref + 2 * e; const real* solimp = m->eq_solimp + 5 * e;
    int o1 = m->eq_obj1id[e], o2 = m->eq_obj2id[e];
    if (type == EQ_TENDON) { /* constraint.py:642-826 */
      w->ne[0] += 1;
      int efcid = nefc; nefc += 1;
      if (efcid >= njmax) continue;
      real* J = w->efc_J + (size_t)efcid * nv, pos, invweight, deriv = 0, Jqvel = 0;
      for (int i = 0; i < nv; i++) J[i] = 0;
      const real pos1 = w->ten_length[o1] - m->tendon_length0[o1];
      if (o2 > -1) {
        invweight = m->tendon_invweight0[o1] + m->tendon_invweight0[o2];
        const real dif = w->ten_length[o2] - m->tendon_length0[o2], dif2 = dif * dif, dif3 = dif2 * dif, dif4 = dif3 * dif;
        pos = pos1 - (data[0] + data[1] * dif + data[2] * dif2 + data[3] * dif3 + data[4] * dif4);
        deriv = data[1] + 2 * data[2] * dif + 3 * data[3] * dif2 + 4 * data[4] * dif3;
      } else {
        invweight = m->tendon_invweight0[o1];
        pos = pos1 - data[0];
      }
      tendon_row(w, o1, 1, J);
      if (deriv != 0) tendon_row(w, o2, -deriv, J);
      for (int i = 0; i < nv; i++) Jqvel += J[i] * w->qvel[i];
      efc_row(w, efcid, pos, pos, invweight, solref, solimp, 0, Jqvel, 0, CNSTR_EQUALITY, e);
      continue;
    }
    if (type == EQ_JOINT) {
      w->ne[0] += 1;
      int efcid = nefc; nefc += 1;
      if (efcid >= njmax) continue;
      int d1 = m->jnt_dofadr[o1], q1 = m->jnt_qposadr[o1];
      real* J = w->efc_J + (size_t)efcid * nv, pos, Jqvel, invweight;
      for (int i = 0; i < nv; i++) J[i] = 0;
      J[d1] = 1;
      if (o2 > -1) {
        int q2 = m->jnt_qposadr[o2], d2 = m->jnt_dofadr[o2];
        real dif = w->qpos[q2] - m->qpos0[q2];
        real rhs = data[0] + dif * (data[1] + dif * (data[2] + dif * (data[3] + dif * data[4])));
        real deriv2 = data[1] + dif * (2 * data[2] + dif * (3 * data[3] + dif * 4 * data[4]));
        pos = w->qpos[q1] - m->qpos0[q1] - rhs;
        Jqvel = w->qvel[d1] - w->qvel[d2] * deriv2;
        invweight = m->dof_invweight0[d1] + m->dof_invweight0[d2];
        J[d2] = -deriv2;
      } else {
        pos = w->qpos[q1] - m->qpos0[q1] - data[0];
        Jqvel = w->qvel[d1];
        invweight = m->dof_invweight0[d1];
      }
      efc_row(w, efcid, pos, pos, invweight, solref, solimp, 0, Jqvel, 0, CNSTR_EQUALITY, e);
      continue;
    }
    const int nrow = type == EQ_CONNECT ? 3 : 6;
    w->ne[0] += nrow;
    int efcid = nefc; nefc += nrow;
    if (efcid >= njmax - nrow) continue;
    const int b1 = o1, b2 = o2;
    const real *a1 = data, *a2 = data + 3; /* connect: anchor1 in body1, anchor2 in body2; weld: data[0:3] lives in body2 */
    real pos1[3], pos2[3], t[3];
    matvec3(w->xmat + 9 * b1, type == EQ_CONNECT ? a1 : a2, t);
    for (int i = 0; i < 3; i++) pos1[i] = w->xpos[3 * b1 + i] + t[i];
    matvec3(w->xmat + 9 * b2, type == EQ_CONNECT ? a2 : a1, t);
    for (int i = 0; i < 3; i++) pos2[i] = w->xpos[3 * b2 + i] + t[i];
    real quat[4] = {1, 0, 0, 0}, quat1[4] = {1, 0, 0, 0}, torquescale = 0;
    if (type == EQ_WELD) {
      torquescale = data[10];
      mul_quat(w->xquat + 4 * b1, data + 6, quat);
      const real* q2 = w->xquat + 4 * b2;
      quat1[0] = q2[0]; quat1[1] = -q2[1]; quat1[2] = -q2[2]; quat1[3] = -q2[3];
    }
    real Jqvelp[3] = {0, 0, 0}, Jqvelr[3] = {0, 0, 0}, Jdotvp[3] = {0, 0, 0}, Jdotvr0[3] = {0, 0, 0};
    for (int d = 0; d < nv; d++) {
      real jp1[3], jr1[3], jp2[3], jr2[3], dp1[3], dr1[3], dp2[3], dr2[3];
      jac_dof(w, pos1, b1, d, jp1, jr1); jac_dof(w, pos2, b2, d, jp2, jr2);
      jac_dot_dof(w, pos1, b1, d, dp1, dr1); jac_dot_dof(w, pos2, b2, d, dp2, dr2);
      real qv = w->qvel[d];
      for (int i = 0; i < 3; i++) {
        real jd = jp1[i] - jp2[i];
        w->efc_J[(size_t)(efcid + i) * nv + d] = jd;
        Jqvelp[i] += jd * qv; Jdotvp[i] += (dp1[i] - dp2[i]) * qv;
      }
      if (type == EQ_WELD) {
        real jdr[3] = {(jr1[0] - jr2[0]) * torquescale, (jr1[1] - jr2[1]) * torquescale, (jr1[2] - jr2[2]) * torquescale}, qa[4], qb[4];
        quat_mul_axis(quat1, jdr, qa);
        mul_quat(qa, quat, qb);
        for (int i = 0; i < 3; i++) {
          real jr = (real)0.5 * qb[1 + i];
          w->efc_J[(size_t)(efcid + 3 + i) * nv + d] = jr;
          Jqvelr[i] += jr * qv; Jdotvr0[i] += (dr1[i] - dr2[i]) * qv;
        }
      }
    }
    real cpos[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
    real invw_t = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
    if (type == EQ_CONNECT) {
      real pos_imp = len3(cpos);
      for (int i = 0; i < 3; i++) { efc_row(w, efcid + i, cpos[i], pos_imp, invw_t, solref, solimp, 0, Jqvelp[i], 0, CNSTR_EQUALITY, e); w->efc_aref[efcid + i] -= Jdotvp[i]; }
      continue;
    }
    real crotq[4], crot[3];
    mul_quat(quat1, quat, crotq);
    for (int i = 0; i < 3; i++) crot[i] = crotq[1 + i] * torquescale;
    real pos_imp = (real)sqrt((double)(dot3(cpos, cpos) + dot3(crot, crot)));
    /* rotational Jdotv from the quaternion product rule (constraint.py:1085-1117, 1381-1395) */
    const real *om1 = w->cvel + 6 * b1, *om2 = w->cvel + 6 * b2;
    real om1q[4] = {0, om1[0], om1[1], om1[2]}, om2q[4] = {0, om2[0], om2[1], om2[2]}, domq[4] = {0, om1[0] - om2[0], om1[1] - om2[1], om1[2] - om2[2]};
    real qdot0[4], qdot0r[4], qdot1[4], negqdot1[4], negq1[4], dj[4] = {0, Jdotvr0[0], Jdotvr0[1], Jdotvr0[2]}, ta[4], t1[4], t2[4], t3[4];
    mul_quat(om1q, w->xquat + 4 * b1, qdot0);
    for (int i = 0; i < 4; i++) qdot0[i] *= (real)0.5;
    mul_quat(qdot0, data + 6, qdot0r);
    mul_quat(om2q, w->xquat + 4 * b2, qdot1);
    for (int i = 0; i < 4; i++) qdot1[i] *= (real)0.5;
    negqdot1[0] = qdot1[0]; negq1[0] = w->xquat[4 * b2];
    for (int i = 1; i < 4; i++) { negqdot1[i] = -qdot1[i]; negq1[i] = -w->xquat[4 * b2 + i]; }
    mul_quat(negqdot1, domq, ta); mul_quat(ta, quat, t1);
    mul_quat(negq1, dj, ta); mul_quat(ta, quat, t2);
    mul_quat(negq1, domq, ta); mul_quat(ta, qdot0r, t3);
    real invw_r = m->body_invweight0[2 * b1 + 1] + m->body_invweight0[2 * b2 + 1];
    for (int i = 0; i < 3; i++) { efc_row(w, efcid + i, cpos[i], pos_imp, invw_t, solref, solimp, 0, Jqvelp[i], 0, CNSTR_EQUALITY, e); w->efc_aref[efcid + i] -= Jdotvp[i]; }
    for (int i = 0; i < 3; i++) {
      efc_row(w, efcid + 3 + i, crot[i], pos_imp, invw_r, solref, solimp, 0, Jqvelr[i], 0, CNSTR_EQUALITY, e);
      w->efc_aref[efcid + 3 + i] -= (t1[1 + i] + t2[1 + i] + t3[1 + i]) * (real)0.5 * torquescale;
    }
  }
  return nefc;
}
static void make_constraint(W* w) {
  const OrcModel* m = w->m;
  const int nv = m->nv, njmax = w->njmax, np = m->nmaxpyramid;
  int nefc = 0;
  w->ne[0] = w->nf[0] = w->nl[0] = 0;
  if (m->disableflags & DSBL_CONSTRAINT) { w->nefc[0] = 0; return; }
  if (!(m->disableflags & DSBL_EQUALITY)) nefc = equality_rows(w, nefc);
  /* dof friction (constraint.py:1765) */
  if (!(m->disableflags & DSBL_FRICTIONLOSS)) {
    for (int d = 0; d < nv; d++) {
      if (m->dof_frictionloss[d] <= 0) continue;
      w->nf[0]++;
      int efcid = nefc++;
      if (efcid >= njmax) continue;
      for (int i = 0; i < nv; i++) w->efc_J[efcid * nv + i] = 0;
      w->efc_J[efcid * nv + d] = 1;
      efc_row(w, efcid, 0, 0, m->dof_invweight0[d], m->dof_solref + 2 * d, m->dof_solimp + 5 * d, 0, w->qvel[d], m->dof_frictionloss[d], CNSTR_FRICTION_DOF, d);
    }
    for (int t = 0; t < m->ntendon; t++) { /* constraint.py:1867-1985 */
      if (m->tendon_frictionloss[t] <= 0) continue;
      w->nf[0]++;
      int efcid = nefc++;
      if (efcid >= njmax) continue;
      real* J = w->efc_J + (size_t)efcid * nv, Jqvel = 0;
      for (int i = 0; i < nv; i++) J[i] = 0;
      tendon_row(w, t, 1, J);
      for (int i = 0; i < nv; i++) Jqvel += J[i] * w->qvel[i];
      efc_row(w, efcid, 0, 0, m->tendon_invweight0[t], m->tendon_solref_fri + 2 * t, m->tendon_solimp_fri + 5 * t, 0, Jqvel, m->tendon_frictionloss[t], CNSTR_FRICTION_TENDON, t);
    }
  }
  /* joint limits: ball (constraint.py:2107) first, then slide/hinge (:1990) -- the reference's launch order */
  if (!(m->disableflags & DSBL_LIMIT)) {
    for (int li = 0; li < m->nlimit_ball; li++) {
      int j = m->jnt_limited_ball_adr[li], qa = m->jnt_qposadr[j];
      real q[4] = {w->qpos[qa], w->qpos[qa + 1], w->qpos[qa + 2], w->qpos[qa + 3]}, axis[3] = {0, 0, 0}, angle = 0;
      normalize4(q);
      real s2 = (real)sqrt((double)(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]));
      if (s2 != 0) { /* math.py:161 quat_to_vel, then normalize_with_norm */
        real speed = 2 * (real)atan2((double)s2, (double)q[0]);
        if (speed > (real)3.14159265358979323846) speed -= 2 * (real)3.14159265358979323846;
        real v[3] = {q[1] * speed / s2, q[2] * speed / s2, q[3] * speed / s2};
        angle = len3(v);
        for (int i = 0; i < 3; i++) axis[i] = angle == 0 ? v[i] : v[i] / angle;
      }
      real margin = m->jnt_margin[j], pos = rmax(m->jnt_range[2 * j], m->jnt_range[2 * j + 1]) - angle - margin;
      if (!(pos < 0)) continue;
      w->nl[0]++;
      int efcid = nefc++;
      if (efcid >= njmax) continue;
      int d = m->jnt_dofadr[j];
      real Jqvel = 0;
      for (int i = 0; i < nv; i++) w->efc_J[efcid * nv + i] = 0;
      for (int i = 0; i < 3; i++) { w->efc_J[efcid * nv + d + i] = -axis[i]; Jqvel -= axis[i] * w->qvel[d + i]; }
      efc_row(w, efcid, pos, pos, m->dof_invweight0[d], m->jnt_solref + 2 * j, m->jnt_solimp + 5 * j, margin, Jqvel, 0, CNSTR_LIMIT_JOINT, j);
    }
    for (int li = 0; li < m->nlimit; li++) {
      int j = m->jnt_limited_slide_hinge_adr[li];
      real qpos = w->qpos[m->jnt_qposadr[j]], margin = m->jnt_margin[j];
      real dist_min = qpos - m->jnt_range[2 * j], dist_max = m->jnt_range[2 * j + 1] - qpos;
      real pos = rmin(dist_min, dist_max) - margin;
      if (!(pos < 0)) continue;
      w->nl[0]++;
      int efcid = nefc++;
      if (efcid >= njmax) continue;
      int d = m->jnt_dofadr[j];
      real J = (dist_min < dist_max ? (real)1 : (real)0) * 2 - 1;
      for (int i = 0; i < nv; i++) w->efc_J[efcid * nv + i] = 0;
      w->efc_J[efcid * nv + d] = J;
      efc_row(w, efcid, pos, pos, m->dof_invweight0[d], m->jnt_solref + 2 * j, m->jnt_solimp + 5 * j, margin, J * w->qvel[d], 0, CNSTR_LIMIT_JOINT, j);
    }
    for (int t = 0; t < m->ntendon; t++) { /* constraint.py:2243-2375 */
      if (!m->tendon_limited[t]) continue;
      const real length = w->ten_length[t], margin = m->tendon_margin[t];
      const real dist_min = length - m->tendon_range[2 * t], dist_max = m->tendon_range[2 * t + 1] - length;
      const real pos = rmin(dist_min, dist_max) - margin;
      if (!(pos < 0)) continue;
      w->nl[0]++;
      int efcid = nefc++;
      if (efcid >= njmax) continue;
      const real scl = (dist_min < dist_max ? (real)1 : (real)0) * 2 - 1;
      real* J = w->efc_J + (size_t)efcid * nv, Jqvel = 0;
      for (int i = 0; i < nv; i++) J[i] = 0;
      tendon_row(w, t, scl, J);
      for (int i = 0; i < nv; i++) Jqvel += J[i] * w->qvel[i];
      efc_row(w, efcid, pos, pos, m->tendon_invweight0[t], m->tendon_solref_lim + 2 * t, m->tendon_solimp_lim + 5 * t, margin, Jqvel, 0, CNSTR_LIMIT_TENDON, t);
    }
  }
  /* contacts (constraint.py:2641 init, :3751 dense jac, :4197 update) */
  if (!(m->disableflags & DSBL_CONTACT)) {
    for (int c = 0; c < w->ncon[0]; c++) {
      int condim = w->con_dim[c];
      real includemargin = w->con_includemargin[c], pos = w->con_dist[c] - includemargin;
      if (!(pos < 0)) continue;
      int ndim = (m->cone == CONE_ELLIPTIC) ? condim : (condim == 1 ? 1 : 2 * (condim - 1));
      int base = nefc; nefc += ndim;
      int g1 = w->con_geom[2 * c], g2 = w->con_geom[2 * c + 1], b1 = m->geom_bodyid[g1], b2 = m->geom_bodyid[g2];
      const real* cpos = w->con_pos + 3 * c; const real* frame = w->con_frame + 9 * c; const real* fri = w->con_friction + 5 * c;
      real off1[3], off2[3];
      for (int i = 0; i < 3; i++) { off1[i] = cpos[i] - w->subtree_com[3 * m->body_rootid[b1] + i]; off2[i] = cpos[i] - w->subtree_com[3 * m->body_rootid[b2] + i]; }
      real invweight0 = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
      for (int dim = 0; dim < ndim; dim++) {
        int efcid = base + dim;
        if (efcid >= njmax) { w->con_efc_address[np * c + dim] = -1; continue; }
        w->con_efc_address[np * c + dim] = efcid;
        real Jqvel = 0;
        for (int d = 0; d < nv; d++) {
          const real* cd = w->cdof + 6 * d;
          real jp1[3] = {0, 0, 0}, jr1[3] = {0, 0, 0}, jp2[3] = {0, 0, 0}, jr2[3] = {0, 0, 0}, cr[3];
          if (m->body_isdofancestor[b1 * nv + d]) { cross3(cd, off1, cr); for (int i = 0; i < 3; i++) { jp1[i] = cd[3 + i] + cr[i]; jr1[i] = cd[i]; } }
          if (m->body_isdofancestor[b2 * nv + d]) { cross3(cd, off2, cr); for (int i = 0; i < 3; i++) { jp2[i] = cd[3 + i] + cr[i]; jr2[i] = cd[i]; } }
          real jpd[3] = {jp2[0] - jp1[0], jp2[1] - jp1[1], jp2[2] - jp1[2]}, jrd[3] = {jr2[0] - jr1[0], jr2[1] - jr1[1], jr2[2] - jr1[2]};
          real J;
          if (m->cone == CONE_ELLIPTIC) {
            J = dim < 3 ? dot3(jpd, frame + 3 * dim) : dot3(jrd, frame + 3 * (dim - 3));
          } else {
            J = dot3(jpd, frame);
            if (condim > 1) {
              int dimid2 = dim / 2 + 1;
              real frii = fri[dimid2 - 1] * (1 - 2 * (real)(dim & 1));
              if (dimid2 == 1) J += dot3(jpd, frame + 3) * frii;
              else if (dimid2 == 2) J += dot3(jpd, frame + 6) * frii;
              else if (dimid2 == 3) J += dot3(jrd, frame) * frii;
              else if (dimid2 == 4) J += dot3(jrd, frame + 3) * frii;
              else J += dot3(jrd, frame + 6) * frii;
            }
          }
          w->efc_J[efcid * nv + d] = J;
          Jqvel += J * w->qvel[d];
        }
        real invweight = invweight0, pos_aref = pos; const real* ref = w->con_solref + 2 * c;
        int type;
        if (m->cone == CONE_ELLIPTIC) {
          if (dim > 0) {
            const real* srf = w->con_solreffriction + 2 * c;
            if (srf[0] != 0 || srf[1] != 0) ref = srf;
            invweight = invweight * m->impratio_invsqrt * m->impratio_invsqrt;
            if (dim > 1) { real f0 = fri[0], fi = fri[dim - 1]; invweight *= f0 * f0 / (fi * fi); }
            pos_aref = 0;
          }
        } else if (condim > 1) {
          real f0 = fri[0];
          invweight = invweight + f0 * f0 * invweight;
          invweight = invweight * 2 * f0 * f0 * m->impratio_invsqrt * m->impratio_invsqrt;
        }
        type = condim == 1 ? CNSTR_CONTACT_FRICTIONLESS : (m->cone == CONE_ELLIPTIC ? CNSTR_CONTACT_ELLIPTIC : CNSTR_CONTACT_PYRAMIDAL);
        efc_row(w, efcid, pos_aref, pos, invweight, ref, w->con_solimp + 5 * c, includemargin, Jqvel, 0, type, c);
      }
    }
  }
  w->nefc[0] = nefc;
}

/* ------------------------------------------------------------------ fwd_velocity (forward.py:732-753) */
static void fwd_velocity(W* w) {
  const OrcModel* m = w->m;
  const int nv = m->nv, nb = m->nbody;
  for (int a = 0; a < m->nu; a++) { /* forward.py:680 */
    real vel = 0;
    for (int i = 0; i < w->moment_rownnz[a]; i++) { int s = w->moment_rowadr[a] + i; vel += w->actuator_moment[s] * w->qvel[w->moment_colind[s]]; }
    w->actuator_velocity[a] = vel;
  }
  for (int t = 0; t < m->ntendon; t++) { /* forward.py:706-729 */
    real vel = 0;
    for (int c = 0; c < m->ten_J_rownnz[t]; c++) { const int s = m->ten_J_rowadr[t] + c; vel += w->ten_J[s] * w->qvel[m->ten_J_colind[s]]; }
    w->ten_velocity[t] = vel;
  }
  /* com_vel (smooth.py:2179-2285) */
  memset(w->cvel, 0, 6 * sizeof(real));
  for (int b = 1; b < nb; b++) {
    int pid = m->body_parentid[b], dofid = m->body_dofadr[b], jntid = m->body_jntadr[b], jntnum = m->body_jntnum[b];
    real cvel[6];
    memcpy(cvel, w->cvel + 6 * pid, sizeof cvel);
    for (int j = jntid; j < jntid + jntnum; j++) {
      int t = m->jnt_type[j];
      if (t == JNT_FREE) {
        for (int k = 0; k < 3; k++) for (int i = 0; i < 6; i++) cvel[i] += w->cdof[6 * (dofid + k) + i] * w->qvel[dofid + k];
        memset(w->cdof_dot + 6 * dofid, 0, 18 * sizeof(real));
        for (int k = 3; k < 6; k++) motion_cross(cvel, w->cdof + 6 * (dofid + k), w->cdof_dot + 6 * (dofid + k));
        for (int k = 3; k < 6; k++) for (int i = 0; i < 6; i++) cvel[i] += w->cdof[6 * (dofid + k) + i] * w->qvel[dofid + k];
        dofid += 6;
      } else if (t == JNT_BALL) {
        for (int k = 0; k < 3; k++) motion_cross(cvel, w->cdof + 6 * (dofid + k), w->cdof_dot + 6 * (dofid + k));
        for (int k = 0; k < 3; k++) for (int i = 0; i < 6; i++) cvel[i] += w->cdof[6 * (dofid + k) + i] * w->qvel[dofid + k];
        dofid += 3;
      } else {
        motion_cross(cvel, w->cdof + 6 * dofid, w->cdof_dot + 6 * dofid);
        for (int i = 0; i < 6; i++) cvel[i] += w->cdof[6 * dofid + i] * w->qvel[dofid];
        dofid += 1;
      }
    }
    memcpy(w->cvel + 6 * b, cvel, sizeof cvel);
  }
  /* passive (passive.py:73-206,631-667,1257): springs/dampers on slide/hinge; free/ball damping; gravcomp/fluid absent */
  int dsbl_spring = m->disableflags & DSBL_SPRING, dsbl_damper = m->disableflags & DSBL_DAMPER;
  for (int d = 0; d < nv; d++) w->qfrc_spring[d] = w->qfrc_damper[d] = w->qfrc_gravcomp[d] = w->qfrc_passive[d] = 0;
  if (!(dsbl_spring && dsbl_damper)) {
    for (int j = 0; j < m->njnt; j++) {
      int d = m->jnt_dofadr[j], t = m->jnt_type[j], qa = m->jnt_qposadr[j];
      real stiffness = m->jnt_stiffness[j];
      int has_st = stiffness != 0 && !dsbl_spring;
      int nd = t == JNT_FREE ? 6 : (t == JNT_BALL ? 3 : 1);
      if (has_st) { /* passive.py:125-206; polynomial stiffness terms are zero on this path */
        if (t == JNT_SLIDE || t == JNT_HINGE) w->qfrc_spring[d] = -(w->qpos[qa] - m->qpos_spring[qa]) * stiffness;
        else {
          int ra = qa, rd = d;
          if (t == JNT_FREE) {
            for (int i = 0; i < 3; i++) w->qfrc_spring[d + i] = -stiffness * (w->qpos[qa + i] - m->qpos_spring[qa + i]);
            ra = qa + 3; rd = d + 3;
          }
          real rot[4] = {w->qpos[ra], w->qpos[ra + 1], w->qpos[ra + 2], w->qpos[ra + 3]}, qneg[4], qdif[4], dif[3] = {0, 0, 0};
          normalize4(rot);
          qneg[0] = m->qpos_spring[ra]; for (int i = 1; i < 4; i++) qneg[i] = -m->qpos_spring[ra + i];
          mul_quat(qneg, rot, qdif); /* math.py:178 quat_sub, :161 quat_to_vel */
          real s2 = (real)sqrt((double)(qdif[1] * qdif[1] + qdif[2] * qdif[2] + qdif[3] * qdif[3]));
          if (s2 != 0) {
            real speed = 2 * (real)atan2((double)s2, (double)qdif[0]);
            if (speed > (real)3.14159265358979323846) speed -= 2 * (real)3.14159265358979323846;
            for (int i = 0; i < 3; i++) dif[i] = qdif[1 + i] * speed / s2;
          }
          for (int i = 0; i < 3; i++) w->qfrc_spring[rd + i] = -stiffness * dif[i];
        }
      }
      for (int k = 0; k < nd; k++) {
        real damping = m->dof_damping[d + k];
        if (damping != 0 && !dsbl_damper) w->qfrc_damper[d + k] = -w->qvel[d + k] * damping;
      }
    }
    for (int t = 0; t < m->ntendon; t++) { /* passive.py:208-272 (polynomial terms zero) */
      const real stiffness = m->tendon_stiffness[t], damping = m->tendon_damping[t];
      if (stiffness != 0 && !dsbl_spring) {
        const real length = w->ten_length[t], lower = m->tendon_lengthspring[2 * t], upper = m->tendon_lengthspring[2 * t + 1];
        const real x = length > upper ? length - upper : (length < lower ? length - lower : 0);
        tendon_row(w, t, -x * stiffness, w->qfrc_spring);
      }
      if (damping != 0 && !dsbl_damper) tendon_row(w, t, -w->ten_velocity[t] * damping, w->qfrc_damper);
    }
    for (int d = 0; d < nv; d++) w->qfrc_passive[d] = w->qfrc_spring[d] + w->qfrc_damper[d];
  }
  /* gravity compensation (passive.py:275-303): -gravity * mass * gravcomp applied at the body's inertial frame origin;
   * added to qfrc_passive unless the joint routes it through the actuators (passive.py:650-653) */
  if (!(m->disableflags & DSBL_GRAVITY) && !(dsbl_spring && dsbl_damper)) { /* passive() returns early when both are disabled (:1263) */
    int any = 0;
    for (int b = 1; b < nb; b++) {
      real gc = m->body_gravcomp[b];
      if (gc == 0) continue;
      any = 1;
      real force[3] = {-m->gravity[0] * m->body_mass[b] * gc, -m->gravity[1] * m->body_mass[b] * gc, -m->gravity[2] * m->body_mass[b] * gc};
      for (int d = 0; d < nv; d++) {
        if (!m->body_isdofancestor[b * nv + d]) continue;
        const real* com = w->subtree_com + 3 * m->body_rootid[b]; const real* cd = w->cdof + 6 * d;
        real off[3] = {w->xipos[3 * b] - com[0], w->xipos[3 * b + 1] - com[1], w->xipos[3 * b + 2] - com[2]}, cr[3];
        cross3(cd, off, cr);
        w->qfrc_gravcomp[d] += (cd[3] + cr[0]) * force[0] + (cd[4] + cr[1]) * force[1] + (cd[5] + cr[2]) * force[2];
      }
    }
    if (any) for (int d = 0; d < nv; d++) if (!m->jnt_actgravcomp[m->dof_jntid[d]]) w->qfrc_passive[d] += w->qfrc_gravcomp[d];
  }
  /* rne (smooth.py:1353-1515), flg_acc = False */
  memset(w->cacc, 0, 6 * sizeof(real));
  if (!(m->disableflags & DSBL_GRAVITY)) for (int i = 0; i < 3; i++) w->cacc[3 + i] = -m->gravity[i];
  for (int b = 1; b < nb; b++) {
    int pid = m->body_parentid[b];
    real a[6];
    memcpy(a, w->cacc + 6 * pid, sizeof a);
    for (int k = 0; k < m->body_dofnum[b]; k++) { int d = m->body_dofadr[b] + k; for (int i = 0; i < 6; i++) a[i] += w->cdof_dot[6 * d + i] * w->qvel[d]; }
    memcpy(w->cacc + 6 * b, a, sizeof a);
  }
  memset(w->cfrc_int, 0, 6 * sizeof(real));
  for (int b = 1; b < nb; b++) {
    real f[6], iv[6], g[6];
    inert_vec(w->cinert + 10 * b, w->cacc + 6 * b, f);
    inert_vec(w->cinert + 10 * b, w->cvel + 6 * b, iv);
    motion_cross_force(w->cvel + 6 * b, iv, g);
    for (int i = 0; i < 6; i++) w->cfrc_int[6 * b + i] = f[i] + g[i];
  }
  for (int b = nb - 1; b >= 1; b--) { int p = m->body_parentid[b]; for (int i = 0; i < 6; i++) w->cfrc_int[6 * p + i] += w->cfrc_int[6 * b + i]; }
  for (int d = 0; d < nv; d++) { real s = 0; int b = m->dof_bodyid[d]; for (int i = 0; i < 6; i++) s += w->cdof[6 * d + i] * w->cfrc_int[6 * b + i]; w->qfrc_bias[d] = s; }
}

/* support.py:38-64 next_act: one integration step of an activation (exact for FILTEREXACT), optionally clamped to actrange */
enum { DYN_NONE = 0, DYN_INTEGRATOR = 1, DYN_FILTER = 2, DYN_FILTEREXACT = 3 };
static real next_act(const OrcModel* m, int a, real act, real act_dot, real scale, int clamp) {
  real r;
  if (m->actuator_dyntype[a] == DYN_FILTEREXACT) {
    real tau = rmax(MJ_MINVAL, m->actuator_dynprm[10 * a]);
    r = act + scale * act_dot * tau * ((real)1 - (real)exp((double)(-m->timestep / tau)));
  } else r = act + scale * act_dot * m->timestep;
  if (clamp) r = rclamp(r, m->actuator_actrange[2 * a], m->actuator_actrange[2 * a + 1]);
  return r;
}
/* forward.py:135-218 _next_activation (INTEGRATOR / FILTER / FILTEREXACT / NONE): act <- next_act(act0, act_dot) */
static void next_activation(W* w, const real* act0, real scale, int limit) {
  const OrcModel* m = w->m;
  for (int a = 0; a < m->nu; a++)
    for (int j = m->actuator_actadr[a]; j >= 0 && j < m->actuator_actadr[a] + m->actuator_actnum[a]; j++)
      w->act[j] = next_act(m, a, act0[j], w->act_dot[j], scale, limit && m->actuator_actlimited[a]);
}

/* ------------------------------------------------------------------ fwd_actuation (forward.py:756-1252; dyntype NONE / INTEGRATOR / FILTER / FILTEREXACT) */
static void fwd_actuation(W* w) {
  const OrcModel* m = w->m;
  for (int d = 0; d < m->nv; d++) w->qfrc_actuator[d] = 0;
  if (!m->nu || (m->disableflags & DSBL_ACTUATION)) {
    for (int a = 0; a < m->nu; a++) w->actuator_force[a] = 0;
    for (int j = 0; j < m->na; j++) w->act_dot[j] = 0;
    return;
  }
  for (int a = 0; a < m->nu; a++) {
    real ctrl = w->ctrl[a];
    if (m->actuator_ctrllimited[a] && !(m->disableflags & DSBL_CLAMPCTRL)) ctrl = rclamp(ctrl, m->actuator_ctrlrange[2 * a], m->actuator_ctrlrange[2 * a + 1]);
    real ctrl_act = ctrl;
    if (m->na && m->actuator_actadr[a] >= 0) { /* forward.py:800-963 */
      const int last = m->actuator_actadr[a] + m->actuator_actnum[a] - 1, dyn = m->actuator_dyntype[a];
      const real act = w->act[last];
      real act_dot = 0;
      if (dyn == DYN_INTEGRATOR) act_dot = ctrl;
      else if (dyn == DYN_FILTER || dyn == DYN_FILTEREXACT) act_dot = (ctrl - act) / rmax(m->actuator_dynprm[10 * a], MJ_MINVAL);
      w->act_dot[last] = act_dot;
      ctrl_act = m->actuator_actearly[a] ? next_act(m, a, act, act_dot, 1, m->actuator_actlimited[a]) : act;
    }
    real length = w->actuator_length[a], velocity = w->actuator_velocity[a];
    const real *gp = m->actuator_gainprm + 10 * a, *bp = m->actuator_biasprm + 10 * a;
    real gain = 0, bias = 0;
    if (m->actuator_gaintype[a] == GAIN_FIXED) gain = gp[0];
    else if (m->actuator_gaintype[a] == GAIN_AFFINE) gain = gp[0] + gp[1] * length + gp[2] * velocity;
    if (m->actuator_biastype[a] == BIAS_AFFINE) bias = bp[0] + bp[1] * length + bp[2] * velocity;
    real force = gain * ctrl_act + bias;
    if (m->actuator_forcelimited[a]) force = rclamp(force, m->actuator_forcerange[2 * a], m->actuator_forcerange[2 * a + 1]);
    w->actuator_force[a] = force;
  }
  for (int t = 0; t < m->ntendon; t++) { /* forward.py:1054-1094: the actuators of a force-limited tendon share its range */
    if (!m->tendon_actfrclimited[t]) continue;
    real total = 0;
    for (int a = 0; a < m->nu; a++) if (m->actuator_trntype[a] == TRN_TENDON && m->actuator_trnid[2 * a] == t) total += w->actuator_force[a];
    const real lo = m->tendon_actfrcrange[2 * t], hi = m->tendon_actfrcrange[2 * t + 1];
    const real sc = total < lo ? lo / total : (total > hi ? hi / total : 1);
    if (sc != 1) for (int a = 0; a < m->nu; a++) if (m->actuator_trntype[a] == TRN_TENDON && m->actuator_trnid[2 * a] == t) w->actuator_force[a] *= sc;
  }
  for (int a = 0; a < m->nu; a++) /* forward.py:1097 */
    for (int i = 0; i < w->moment_rownnz[a]; i++) { int s = w->moment_rowadr[a] + i; w->qfrc_actuator[w->moment_colind[s]] += w->actuator_moment[s] * w->actuator_force[a]; }
  int gravity_enabled = !(m->disableflags & DSBL_GRAVITY);
  for (int d = 0; d < m->nv; d++) { /* forward.py:1120 */
    int j = m->dof_jntid[d];
    real q = w->qfrc_actuator[d];
    if (gravity_enabled && m->jnt_actgravcomp[j]) q += w->qfrc_gravcomp[d];
    if (m->jnt_actfrclimited[j]) q = rclamp(q, m->jnt_actfrcrange[2 * j], m->jnt_actfrcrange[2 * j + 1]);
    w->qfrc_actuator[d] = q;
  }
}

/* ------------------------------------------------------------------ fwd_acceleration (forward.py:1255-1324, support.py:259-324) */
static void fwd_acceleration(W* w) {
  const OrcModel* m = w->m;
  const int nv = m->nv, nb = m->nbody;
  for (int d = 0; d < nv; d++) w->qfrc_smooth[d] = w->qfrc_passive[d] - w->qfrc_bias[d] + w->qfrc_actuator[d] + w->qfrc_applied[d];
  for (int d = 0; d < nv; d++) { /* _apply_ft */
    const real* cd = w->cdof + 6 * d;
    int db = m->dof_bodyid[d];
    real acc = 0;
    for (int b = db; b < nb; b++) {
      const real* ft = w->xfrc_applied + 6 * b;
      if (ft[0] == 0 && ft[1] == 0 && ft[2] == 0 && ft[3] == 0 && ft[4] == 0 && ft[5] == 0) continue;
      int p = b;
      while (p != 0 && p != db) p = m->body_parentid[p];
      if (p == 0) continue;
      real off[3], cr[3];
      for (int i = 0; i < 3; i++) off[i] = w->xipos[3 * b + i] - w->subtree_com[3 * m->body_rootid[b] + i];
      cross3(cd, off, cr);
      acc += cd[3] * ft[0] + cd[4] * ft[1] + cd[5] * ft[2] + cd[0] * ft[3] + cd[1] * ft[4] + cd[2] * ft[5] + dot3(cr, ft);
    }
    w->qfrc_smooth[d] += acc;
  }
  factor_solve_i(w, w->M, NULL, w->qLD, w->qacc_smooth, w->qfrc_smooth);
}

/* ------------------------------------------------------------------ solver (solver.py) -- Newton, full rebuild path
 * (the reference's incremental-H / stable-state fast path, solver.py:1951,2145-2159, is an algebraically
 * equivalent optimisation of this exact iteration) */
typedef struct {
  real *Jaref, *jv, *grad, *search, *mv, *H, *Hf, *tmp, *Mgrad, *prev_grad, *prev_Mgrad;
  real search_dot, grad_dot, newton_decrement, improvement;
  /* elliptic cones: per contact quad (3), quad1 (u0, v0, uu), quad2 (uv, vv, dm) -- solver.py:957-1015 */
  real* quad;
} SCtx;

/* solver.py:425-477 (pyramidal / frictionless / limit / friction / equality rows) */
static void eval_constraint(int is_equality, int is_friction, real jaref, real D, real frictionloss, real* force, int* state) {
  if (is_equality) { *force = -D * jaref; *state = ST_QUADRATIC; return; }
  if (is_friction) {
    real rf = safe_div(frictionloss, D);
    if (jaref <= -rf) { *force = frictionloss; *state = ST_LINEARNEG; }
    else if (jaref >= rf) { *force = -frictionloss; *state = ST_LINEARPOS; }
    else { *force = -D * jaref; *state = ST_QUADRATIC; }
    return;
  }
  if (jaref >= 0) { *force = 0; *state = ST_SATISFIED; } else { *force = -D * jaref; *state = ST_QUADRATIC; }
}
/* _update_constraint (solver.py:1698-1822,1912-1948); elliptic rows follow _eval_constraint :425-477 and _eval_elliptic_middle :407-422 */
static void update_constraint(W* w, SCtx* c, int nefc) {
  const OrcModel* m = w->m; const int nv = m->nv, ne = w->ne[0], nf = w->nf[0], np = m->nmaxpyramid;
  for (int e = 0; e < nefc; e++) {
    if (w->efc_type[e] == CNSTR_CONTACT_ELLIPTIC) continue; /* handled per contact below */
    int is_eq = e < ne, is_fr = !is_eq && e < ne + nf;
    eval_constraint(is_eq, is_fr, c->Jaref[e], w->efc_D[e], is_fr ? w->efc_frictionloss[e] : 0, &w->efc_force[e], &w->efc_state[e]);
  }
  if (m->cone == CONE_ELLIPTIC) {
    for (int k = 0; k < w->ncon[0]; k++) {
      int dim = w->con_dim[k], e0 = w->con_efc_address[np * k];
      if (dim == 1 || e0 < 0 || e0 >= nefc) continue;
      const real* fri = w->con_friction + 5 * k;
      real mu = fri[0] * m->impratio_invsqrt, N = c->Jaref[e0] * mu, TT = 0;
      for (int j = 1; j < dim; j++) { int ej = w->con_efc_address[np * k + j]; real uj = c->Jaref[ej] * fri[j - 1]; TT += uj * uj; }
      real T = TT <= 0 ? 0 : (real)sqrt((double)TT);
      for (int j = 0; j < dim; j++) {
        int e = w->con_efc_address[np * k + j];
        real jaref = c->Jaref[e], D = w->efc_D[e];
        if ((N >= mu * T) || (T <= 0 && N >= 0)) { w->efc_force[e] = 0; w->efc_state[e] = ST_SATISFIED; }
        else if ((mu * N + T <= 0) || (T <= 0 && N < 0)) { w->efc_force[e] = -D * jaref; w->efc_state[e] = ST_QUADRATIC; }
        else {
          real dm = safe_div(w->efc_D[e0], mu * mu * (1 + mu * mu)), nmt = N - mu * T, fn = -dm * nmt * mu;
          if (j == 0) w->efc_force[e] = fn;
          else { real uf = jaref * fri[j - 1] * fri[j - 1]; w->efc_force[e] = -safe_div(fn, T) * uf; }
          w->efc_state[e] = ST_CONE;
        }
      }
    }
  }
  for (int d = 0; d < nv; d++) { real s = 0; for (int e = 0; e < nefc; e++) s += w->efc_J[e * nv + d] * w->efc_force[e]; w->qfrc_constraint[d] = s; }
}
/* _update_gradient (solver.py:3061-3200): grad, H = M + J^T D_active J, Cholesky, search */
static void update_gradient(W* w, SCtx* c, int nefc) {
  const OrcModel* m = w->m; const int nv = m->nv;
  c->grad_dot = 0;
  for (int d = 0; d < nv; d++) { real g = w->efc_Ma[d] - w->qfrc_smooth[d] - w->qfrc_constraint[d]; c->grad[d] = g; c->grad_dot += g * g; }
  if (m->solver == SOL_CG) { /* Mgrad = M^-1 grad through the factor of M (smooth.solve_m, solver.py:3086) */
    for (int t = 0; t < m->ntree; t++) {
      int start = m->tree_dofadr[t], size = m->tree_dofnum[t];
      chol_upper_solve(w->qLD + m->qLD_block_adr[start], size, c->grad + start, c->Mgrad + start);
    }
    return;
  }
  memset(c->H, 0, (size_t)nv * nv * sizeof(real));
  for (int i = 0; i < nv; i++) { /* densify M (upper + lower) */
    int adr = m->M_rowadr[i];
    for (int k = 0; k < m->M_rownnz[i]; k++) { int j = m->M_colind[adr + k]; c->H[i * nv + j] = w->M[adr + k]; c->H[j * nv + i] = w->M[adr + k]; }
  }
  for (int e = 0; e < nefc; e++) {
    if (w->efc_state[e] != ST_QUADRATIC) continue;
    real D = w->efc_D[e]; const real* J = w->efc_J + e * nv;
    for (int i = 0; i < nv; i++) { if (J[i] == 0) continue; real di = D * J[i]; for (int j = 0; j < nv; j++) c->H[i * nv + j] += di * J[j]; }
  }
  if (m->cone == CONE_ELLIPTIC) { /* JTCJ, solver.py:2443-2565 */
    const int np = m->nmaxpyramid;
    for (int k = 0; k < w->ncon[0]; k++) {
      int dim = w->con_dim[k], e0 = w->con_efc_address[np * k];
      if (dim == 1 || e0 < 0 || e0 >= nefc || w->efc_state[e0] != ST_CONE) continue;
      const real* fri = w->con_friction + 5 * k;
      real mu = fri[0] * m->impratio_invsqrt, mu2 = mu * mu, dm = safe_div(w->efc_D[e0], mu2 * (1 + mu2));
      if (dm == 0) continue;
      real n = c->Jaref[e0] * mu, tt = 0;
      for (int j = 1; j < dim; j++) { int ej = w->con_efc_address[np * k + j]; real u = c->Jaref[ej] * fri[j - 1]; tt += u * u; }
      real t = rmax((real)sqrt((double)tt), MJ_MINVAL), ttt = rmax(t * t * t, MJ_MINVAL), mu_tinv = safe_div(mu, t);
      real mu_n_over_ttt = mu * safe_div(n, ttt), tangent_diag = mu2 - n * mu_tinv;
      for (int d1 = 0; d1 < nv; d1++) for (int d2 = 0; d2 < nv; d2++) {
        real z01 = mu * w->efc_J[e0 * nv + d1], z02 = mu * w->efc_J[e0 * nv + d2], p1 = 0, p2 = 0, td = 0;
        for (int j = 1; j < dim; j++) {
          int ej = w->con_efc_address[np * k + j]; real sc = fri[j - 1], u = c->Jaref[ej] * sc;
          real z1 = sc * w->efc_J[ej * nv + d1], z2 = sc * w->efc_J[ej * nv + d2];
          p1 += u * z1; p2 += u * z2; td += z1 * z2;
        }
        c->H[d1 * nv + d2] += dm * (z01 * z02 - mu_tinv * (z01 * p2 + z02 * p1) + mu_n_over_ttt * p1 * p2 + tangent_diag * td);
      }
    }
  }
  memcpy(c->Hf, c->H, (size_t)nv * nv * sizeof(real));
  chol_upper(c->Hf, nv);
  chol_upper_solve(c->Hf, nv, c->grad, c->tmp);
  c->search_dot = 0; c->newton_decrement = 0;
  for (int d = 0; d < nv; d++) { c->search_dot += c->tmp[d] * c->tmp[d]; c->newton_decrement += c->grad[d] * c->tmp[d]; c->search[d] = -c->tmp[d]; }
}
/* (cost(alpha)-cost(0), grad, hess) of one pyramidal row: solver.py:479-517 */
static void eval_pt_row(int efcid, real alpha, int ne, int nf, real D, real frictionloss, real jaref, real jv, real out[3]) {
  if (efcid >= ne + nf) {
    real x = jaref + alpha * jv, quad0 = (real)0.5 * D * jaref * jaref, cost0 = jaref < 0 ? quad0 : 0, offset = quad0 - cost0;
    if (x < 0) { real jvD = jv * D, h = jv * jvD, ah = alpha * h; out[0] = alpha * (jvD * jaref + (real)0.5 * ah) + offset; out[1] = jvD * jaref + ah; out[2] = h; }
    else { out[0] = -cost0; out[1] = 0; out[2] = 0; }
    return;
  }
  if (efcid >= ne) {
    real f = frictionloss, x = jaref + alpha * jv, rf = safe_div(f, D), c0, p[3];
    if (-rf < jaref && jaref < rf) c0 = (real)0.5 * D * jaref * jaref; else if (jaref <= -rf) c0 = f * ((real)-0.5 * rf - jaref); else c0 = f * ((real)-0.5 * rf + jaref);
    if (-rf < x && x < rf) { real jvD = jv * D; p[0] = (real)0.5 * D * x * x; p[1] = jvD * x; p[2] = jv * jvD; }
    else if (x <= -rf) { p[0] = f * ((real)-0.5 * rf - x); p[1] = -f * jv; p[2] = 0; }
    else { p[0] = f * ((real)-0.5 * rf + x); p[1] = f * jv; p[2] = 0; }
    out[0] = p[0] - c0; out[1] = p[1]; out[2] = p[2];
    return;
  }
  { real jvD = jv * D, h = jv * jvD, ah = alpha * h; out[0] = alpha * (jvD * jaref + (real)0.5 * ah); out[1] = jvD * jaref + ah; out[2] = h; }
}
/* alpha = 0 variant with absolute cost (solver.py:570-592) */
static void eval_pt_row_zero(int efcid, int ne, int nf, real D, real frictionloss, real jaref, real jv, real out[3]) {
  out[0] = out[1] = out[2] = 0;
  if (efcid >= ne + nf) { if (jaref < 0) { real jvD = jv * D; out[0] = (real)0.5 * D * jaref * jaref; out[1] = jvD * jaref; out[2] = jv * jvD; } return; }
  if (efcid >= ne) {
    real f = frictionloss, rf = safe_div(f, D), x = jaref;
    if (-rf < x && x < rf) { real jvD = jv * D; out[0] = (real)0.5 * D * x * x; out[1] = jvD * x; out[2] = jv * jvD; }
    else if (x <= -rf) { out[0] = f * ((real)-0.5 * rf - x); out[1] = -f * jv; out[2] = 0; }
    else { out[0] = f * ((real)-0.5 * rf + x); out[1] = f * jv; out[2] = 0; }
    return;
  }
  { real jvD = jv * D; out[0] = (real)0.5 * D * jaref * jaref; out[1] = jvD * jaref; out[2] = jv * jvD; }
}
/* _eval_elliptic_reference (solver.py:286-305): cost, T, r and state of a contact at alpha = 0 */
static void ell_reference(real mu, const real* q, const real* q1, const real* q2, real* cost0, real* T0, real* r0, int* state0) {
  real u0 = q1[0], uu = q1[2], dm = q2[2];
  if (uu <= 0) { *T0 = 0; *r0 = 0; if (u0 < 0) { *cost0 = q[0]; *state0 = ST_QUADRATIC; } else { *cost0 = 0; *state0 = ST_SATISFIED; } return; }
  *T0 = (real)sqrt((double)uu);
  if (u0 >= mu * *T0) { *cost0 = 0; *r0 = 0; *state0 = ST_SATISFIED; return; }
  if (mu * u0 + *T0 <= 0) { *cost0 = q[0]; *r0 = 0; *state0 = ST_QUADRATIC; return; }
  *r0 = u0 - mu * *T0; *cost0 = (real)0.5 * dm * *r0 * *r0; *state0 = ST_CONE;
}
/* quadratic-zone value shifted by the alpha = 0 cost (solver.py:329-350) */
static void ell_quadratic_shifted(real mu, const real* q, real alpha, real N, real Tsqr, real u0, real T0, real dm, int state0, real out[3]) {
  real aq2 = alpha * q[2], cost = alpha * (aq2 + q[1]);
  if (state0 == ST_CONE) { real b = mu * u0 + T0; cost += (real)0.5 * dm * b * b; }
  else if (state0 == ST_SATISFIED) cost = (real)0.5 * dm * (1 + mu * mu) * (N * N + rmax(Tsqr, 0));
  out[0] = cost; out[1] = 2 * aq2 + q[1]; out[2] = 2 * q[2];
}
/* _eval_elliptic_shifted (solver.py:353-404): (cost(alpha) - cost(0), grad, hess) of one elliptic contact */
static void ell_shifted(real mu, const real* q, const real* q1, const real* q2, real alpha, real cost0, real T0, real r0, int state0, real out[3]) {
  real u0 = q1[0], v0 = q1[1], uu = q1[2], uv = q2[0], vv = q2[1], dm = q2[2];
  real N = u0 + alpha * v0, Tsqr_delta = alpha * (2 * uv + alpha * vv), Tsqr = uu + Tsqr_delta;
  if (Tsqr <= 0) {
    if (N < 0) { ell_quadratic_shifted(mu, q, alpha, N, Tsqr, u0, T0, dm, state0, out); return; }
  } else {
    real T = (real)sqrt((double)Tsqr);
    if (N >= mu * T) { /* top zone: satisfied */ }
    else if (mu * N + T <= 0) { ell_quadratic_shifted(mu, q, alpha, N, Tsqr, u0, T0, dm, state0, out); return; }
    else {
      real T1 = (uv + alpha * vv) / T, T2 = (vv - T1 * T1) / T, r = N - mu * T, r1 = v0 - mu * T1, cost;
      if (state0 == ST_CONE) { real Td = Tsqr_delta / (T + T0), rd = alpha * v0 - mu * Td; cost = (real)0.5 * dm * rd * (2 * r0 + rd); }
      else if (state0 == ST_QUADRATIC) { real aq2 = alpha * q[2], b = mu * N + T; cost = alpha * (aq2 + q[1]) - (real)0.5 * dm * b * b; }
      else cost = (real)0.5 * dm * r * r;
      out[0] = cost; out[1] = dm * r * r1; out[2] = dm * (r1 * r1 + r * (-mu * T2));
      return;
    }
  }
  out[0] = -cost0; out[1] = 0; out[2] = 0;
}
/* absolute value at alpha = 0 (solver.py:308-320) */
static void ell_alpha_zero(real mu, const real* q, const real* q1, const real* q2, real out[3]) {
  real cost0, T0, r0; int st;
  ell_reference(mu, q, q1, q2, &cost0, &T0, &r0, &st);
  out[0] = out[1] = out[2] = 0;
  if (st == ST_QUADRATIC) { out[0] = q[0]; out[1] = q[1]; out[2] = 2 * q[2]; }
  else if (st == ST_CONE) {
    real T1 = q2[0] / T0, T2 = (q2[1] - T1 * T1) / T0, r1 = q1[1] - mu * T1, dm = q2[2];
    out[0] = cost0; out[1] = dm * r0 * r1; out[2] = dm * (r1 * r1 - mu * r0 * T2);
  }
}
/* exported for the reference's known-answer vectors (solver_test.py:296-350): _compute_efc_eval_pt_elliptic on a primary row */
void orc_elliptic_eval_pt(double alpha, const real* quad, const real* quad1, const real* quad2, double mu, real* out) {
  real cost0, T0, r0; int st;
  ell_reference((real)mu, quad, quad1, quad2, &cost0, &T0, &r0, &st);
  ell_shifted((real)mu, quad, quad1, quad2, (real)alpha, cost0, T0, r0, st, out);
}
void orc_elliptic_zero(const real* quad, const real* quad1, const real* quad2, double mu, real* out) { ell_alpha_zero((real)mu, quad, quad1, quad2, out); }
/* per-contact quad / quad1 / quad2 for the current (Jaref, jv) -- solver.py:957-1015 */
static void ell_prepare(W* w, SCtx* c, int nefc) {
  const OrcModel* m = w->m; const int np = m->nmaxpyramid;
  for (int k = 0; k < w->ncon[0]; k++) {
    real* Q = c->quad + 9 * k;
    for (int i = 0; i < 9; i++) Q[i] = 0;
    int dim = w->con_dim[k], e0 = w->con_efc_address[np * k];
    if (dim == 1 || e0 < 0 || e0 >= nefc) continue;
    const real* fri = w->con_friction + 5 * k;
    real mu = fri[0] * m->impratio_invsqrt, ja = c->Jaref[e0], jv = c->jv[e0], D = w->efc_D[e0], jvD = jv * D;
    Q[0] = (real)0.5 * ja * ja * D; Q[1] = jvD * ja; Q[2] = (real)0.5 * jv * jvD;
    real uu = 0, uv = 0, vv = 0;
    for (int j = 1; j < dim; j++) {
      int ej = w->con_efc_address[np * k + j];
      real jvj = c->jv[ej], jaj = c->Jaref[ej], dj = w->efc_D[ej], DJ = dj * jaj;
      Q[0] += (real)0.5 * jaj * DJ; Q[1] += jvj * DJ; Q[2] += (real)0.5 * jvj * dj * jvj;
      real uj = jaj * fri[j - 1], vj = jvj * fri[j - 1];
      uu += uj * uj; uv += uj * vj; vv += vj * vj;
    }
    Q[3] = ja * mu; Q[4] = jv * mu; Q[5] = uu;
    Q[6] = uv; Q[7] = vv; Q[8] = D / (mu * mu * (1 + mu * mu));
  }
}
static void ell_total(W* w, SCtx* c, int nefc, real alpha, int zero, real out[3]) {
  const OrcModel* m = w->m; const int np = m->nmaxpyramid;
  for (int k = 0; k < w->ncon[0]; k++) {
    int dim = w->con_dim[k], e0 = w->con_efc_address[np * k];
    if (dim == 1 || e0 < 0 || e0 >= nefc) continue;
    const real* Q = c->quad + 9 * k; real r[3], mu = w->con_friction[5 * k] * m->impratio_invsqrt;
    if (zero) ell_alpha_zero(mu, Q, Q + 3, Q + 6, r);
    else { real c0, T0, r0; int st; ell_reference(mu, Q, Q + 3, Q + 6, &c0, &T0, &r0, &st); ell_shifted(mu, Q, Q + 3, Q + 6, alpha, c0, T0, r0, st, r); }
    out[0] += r[0]; out[1] += r[1]; out[2] += r[2];
  }
}
static void eval_total(W* w, SCtx* c, int nefc, real alpha, const real quad_gauss[3], real out[3]) {
  /* _eval_pt(quad_gauss, alpha) + sum of rows (solver.py:203-211,1113-1169) */
  const int ne = w->ne[0], nf = w->nf[0];
  real aq2 = alpha * quad_gauss[2];
  out[0] = alpha * aq2 + alpha * quad_gauss[1] + quad_gauss[0]; out[1] = 2 * aq2 + quad_gauss[1]; out[2] = 2 * quad_gauss[2];
  for (int e = 0; e < nefc; e++) {
    if (w->efc_type[e] == CNSTR_CONTACT_ELLIPTIC) continue;
    real r[3]; eval_pt_row(e, alpha, ne, nf, w->efc_D[e], w->efc_frictionloss[e], c->Jaref[e], c->jv[e], r); out[0] += r[0]; out[1] += r[1]; out[2] += r[2];
  }
  if (w->m->cone == CONE_ELLIPTIC) ell_total(w, c, nefc, alpha, 0, out);
}
static int in_bracket(const real* x, const real* y) { return (x[1] < y[1] && y[1] < 0) || (x[1] > y[1] && y[1] > 0); }
#define CP3(d, s) do { (d)[0] = (s)[0]; (d)[1] = (s)[1]; (d)[2] = (s)[2]; } while (0)
/* _linesearch (solver.py:836-1347, pyramidal path) */
static void linesearch(W* w, SCtx* c, int nefc) {
  const OrcModel* m = w->m; const int nv = m->nv, ne = w->ne[0], nf = w->nf[0];
  mul_m(m, w->M, c->search, c->mv);
  for (int e = 0; e < nefc; e++) { real s = 0; for (int d = 0; d < nv; d++) s += w->efc_J[e * nv + d] * c->search[d]; c->jv[e] = s; }
  real snorm = (real)sqrt((double)c->search_dot), scale = m->meaninertia * (real)nv;
  real gtol = rmax(m->tolerance * m->ls_tolerance * snorm * scale, (real)1e-6);
  real p0s[3] = {0, 0, 0};
  if (m->cone == CONE_ELLIPTIC) ell_prepare(w, c, nefc);
  for (int e = 0; e < nefc; e++) {
    if (w->efc_type[e] == CNSTR_CONTACT_ELLIPTIC) continue;
    real r[3]; eval_pt_row_zero(e, ne, nf, w->efc_D[e], w->efc_frictionloss[e], c->Jaref[e], c->jv[e], r); p0s[0] += r[0]; p0s[1] += r[1]; p0s[2] += r[2];
  }
  if (m->cone == CONE_ELLIPTIC) ell_total(w, c, nefc, 0, 1, p0s);
  real qg[3] = {0, 0, 0};
  for (int d = 0; d < nv; d++) { qg[1] += c->search[d] * (w->efc_Ma[d] - w->qfrc_smooth[d]); qg[2] += (real)0.5 * c->search[d] * c->mv[d]; }
  real p0[3] = {qg[0] + p0s[0], qg[1] + p0s[1], 2 * qg[2] + p0s[2]};
  real p0_delta[3] = {0, p0[1], p0[2]};
  real lo_alpha_in = -safe_div(p0[1], p0[2]);
  real lo_in[3];
  eval_total(w, c, nefc, lo_alpha_in, qg, lo_in);
  int initial_converged = fabs((double)lo_in[1]) < gtol && lo_in[0] < 0;
  int ls_converged = initial_converged;
  real alpha = 0, improvement = 0;
  if (!initial_converged) {
    int lo_less = lo_in[1] < p0[1];
    real lo[3], hi[3], lo_alpha, hi_alpha;
    if (lo_less) { CP3(lo, lo_in); lo_alpha = lo_alpha_in; CP3(hi, p0_delta); hi_alpha = 0; }
    else { CP3(lo, p0_delta); lo_alpha = 0; CP3(hi, lo_in); hi_alpha = lo_alpha_in; }
    for (int it = 0; it < m->ls_iterations; it++) {
      real lo_next_alpha = lo_alpha - safe_div(lo[1], lo[2]), hi_next_alpha = hi_alpha - safe_div(hi[1], hi[2]), mid_alpha = (real)0.5 * (lo_alpha + hi_alpha);
      real lo_next[3], hi_next[3], mid[3];
      eval_total(w, c, nefc, lo_next_alpha, qg, lo_next);
      eval_total(w, c, nefc, hi_next_alpha, qg, hi_next);
      eval_total(w, c, nefc, mid_alpha, qg, mid);
      int s1 = in_bracket(lo, lo_next); if (s1) { CP3(lo, lo_next); lo_alpha = lo_next_alpha; }
      int s2 = in_bracket(lo, mid); if (s2) { CP3(lo, mid); lo_alpha = mid_alpha; }
      int s3 = in_bracket(lo, hi_next); if (s3) { CP3(lo, hi_next); lo_alpha = hi_next_alpha; }
      int swap_lo = s1 || s2 || s3;
      int h1 = in_bracket(hi, hi_next); if (h1) { CP3(hi, hi_next); hi_alpha = hi_next_alpha; }
      int h2 = in_bracket(hi, mid); if (h2) { CP3(hi, mid); hi_alpha = mid_alpha; }
      int h3 = in_bracket(hi, lo_next); if (h3) { CP3(hi, lo_next); hi_alpha = lo_next_alpha; }
      int swap_hi = h1 || h2 || h3;
      int ls_done = (!swap_lo && !swap_hi) || (lo[0] < 0 && lo[1] < 0 && lo[1] > -gtol) || (hi[0] < 0 && hi[1] > 0 && hi[1] < gtol);
      int improved = lo[0] < 0 || hi[0] < 0, lo_better = lo[0] < hi[0];
      if (improved) { alpha = lo_better ? lo_alpha : hi_alpha; improvement = -(lo_better ? lo[0] : hi[0]); }
      if (ls_done) { ls_converged = 1; break; }
    }
  } else { alpha = lo_alpha_in; improvement = -lo_in[0]; }
  for (int d = 0; d < nv; d++) { w->qacc[d] += alpha * c->search[d]; w->efc_Ma[d] += alpha * c->mv[d]; }
  for (int e = 0; e < nefc; e++) c->Jaref[e] += alpha * c->jv[e];
  c->improvement = improvement;
  if (!ls_converged) w->overflow[0] |= OVF_LS_ITERATIONS;
}
/* solve (solver.py:3671-3743) */
static void solve(W* w) {
  const OrcModel* m = w->m; const int nv = m->nv;
  if (w->njmax == 0 || nv == 0) { memcpy(w->qacc, w->qacc_smooth, nv * sizeof(real)); w->solver_niter[0] = 0; return; }
  int nefc = w->nefc[0] < w->njmax ? w->nefc[0] : w->njmax;
  SCtx c; memset(&c, 0, sizeof c);
  size_t nr = (size_t)(nefc > 0 ? nefc : 1);
  real* buf = (real*)calloc(2 * nr + 7 * (size_t)nv + 2 * (size_t)nv * nv + 9 * (size_t)(w->nconmax + 1), sizeof(real));
  c.quad = buf + 2 * nr + 4 * (size_t)nv + 2 * (size_t)nv * nv;
  c.Mgrad = c.quad + 9 * (size_t)(w->nconmax + 1); c.prev_grad = c.Mgrad + nv; c.prev_Mgrad = c.prev_grad + nv;
  c.Jaref = buf; c.jv = c.Jaref + nr; c.grad = c.jv + nr; c.search = c.grad + nv; c.mv = c.search + nv; c.tmp = c.mv + nv; c.H = c.tmp + nv; c.Hf = c.H + (size_t)nv * nv;
  const real* start = (m->disableflags & DSBL_WARMSTART) ? w->qacc_smooth : w->qacc_warmstart;
  memcpy(w->qacc, start, nv * sizeof(real));
  w->solver_niter[0] = 0;
  for (int e = 0; e < nefc; e++) { real s = 0; for (int d = 0; d < nv; d++) s += w->efc_J[e * nv + d] * w->qacc[d]; c.Jaref[e] = s - w->efc_aref[e]; }
  mul_m(m, w->M, w->qacc, w->efc_Ma);
  update_constraint(w, &c, nefc);
  update_gradient(w, &c, nefc);
  const int cg = m->solver == SOL_CG;
  if (cg) { /* _solve_init_search_cg (solver.py:1665): search = -Mgrad */
    c.search_dot = 0;
    for (int d = 0; d < nv; d++) { c.search[d] = -c.Mgrad[d]; c.search_dot += c.search[d] * c.search[d]; c.prev_grad[d] = c.grad[d]; c.prev_Mgrad[d] = c.Mgrad[d]; }
  }
  real scale = m->meaninertia * (real)nv;
  int done = m->iterations == 0;
  while (!done) {
    linesearch(w, &c, nefc);
    update_constraint(w, &c, nefc);
    update_gradient(w, &c, nefc);
    w->solver_niter[0] += 1; /* _solve_done solver.py:3453 / _solve_cg_finalize :3402 */
    real improvement = c.improvement / scale, gradient = (real)sqrt((double)c.grad_dot) / scale, model_improvement = (real)0.5 * c.newton_decrement / scale;
    if (cg) { /* Polak-Ribiere (solver.py:3295-3333, 3360-3398) */
      real num = 0, den = 0;
      for (int d = 0; d < nv; d++) { num += c.grad[d] * (c.Mgrad[d] - c.prev_Mgrad[d]); den += c.prev_grad[d] * c.prev_Mgrad[d]; }
      real beta = rmax(0, num / rmax(MJ_MINVAL, den));
      done = improvement < m->tolerance || gradient < m->tolerance;
      if (!done && w->solver_niter[0] == m->iterations) { w->overflow[0] |= OVF_ITERATIONS; done = 1; }
      if (!done) {
        c.search_dot = 0;
        for (int d = 0; d < nv; d++) { c.search[d] = -c.Mgrad[d] + beta * c.search[d]; c.search_dot += c.search[d] * c.search[d]; c.prev_grad[d] = c.grad[d]; c.prev_Mgrad[d] = c.Mgrad[d]; }
      }
      continue;
    }
    done = improvement < m->tolerance || gradient < m->tolerance || model_improvement < m->tolerance;
    if (!done && w->solver_niter[0] == m->iterations) { w->overflow[0] |= OVF_ITERATIONS; done = 1; }
  }
  free(buf);
}

/* ------------------------------------------------------------------ integrators (forward.py:53-131,221-349,387-417) */
/* forward.py:53-115 _next_position: qpos <- integrate(qpos_in, scale * qvel) over one timestep */
static void next_position(const OrcModel* m, const real* qpos_in, const real* qvel, real scale, real* qpos) {
  for (int j = 0; j < m->njnt; j++) {
    int t = m->jnt_type[j], qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (t == JNT_FREE) {
      real ang[3] = {qvel[da + 3] * scale, qvel[da + 4] * scale, qvel[da + 5] * scale};
      for (int i = 0; i < 3; i++) qpos[qa + i] = qpos_in[qa + i] + m->timestep * (qvel[da + i] * scale);
      for (int i = 3; i < 7; i++) qpos[qa + i] = qpos_in[qa + i];
      quat_integrate(qpos + qa + 3, ang, m->timestep);
    } else if (t == JNT_BALL) {
      real ang[3] = {qvel[da] * scale, qvel[da + 1] * scale, qvel[da + 2] * scale};
      for (int i = 0; i < 4; i++) qpos[qa + i] = qpos_in[qa + i];
      quat_integrate(qpos + qa, ang, m->timestep);
    } else {
      qpos[qa] = qpos_in[qa] + m->timestep * qvel[da] * scale;
    }
  }
}
/* forward.py:276-349 _advance: velocity from qacc, position from qvel_pos (the new velocity when NULL: semi-implicit) */
static void advance2(W* w, const real* qacc, const real* qvel_pos) {
  const OrcModel* m = w->m;
  if (m->na) next_activation(w, w->act, 1, 1); /* forward.py:280-300 */
  for (int d = 0; d < m->nv; d++) w->qvel[d] += qacc[d] * m->timestep;
  next_position(m, w->qpos, qvel_pos ? qvel_pos : w->qvel, 1, w->qpos);
  w->time[0] += m->timestep;
  if (w->nefc[0] > w->njmax) w->overflow[0] |= OVF_NEFC;
  memcpy(w->qacc_warmstart, w->qacc, m->nv * sizeof(real));
}
static void advance(W* w, const real* qacc) { advance2(w, qacc, NULL); }
static void euler(W* w) {
  const OrcModel* m = w->m;
  if (!(m->disableflags & (DSBL_EULERDAMP | DSBL_DAMPER))) {
    int qld = 0; for (int t = 0; t < m->ntree; t++) qld += m->tree_dofnum[t] * m->tree_dofnum[t];
    real* L = (real*)malloc(((size_t)qld + 2 * m->nv) * sizeof(real));
    real* qacc = L + qld; real* dd = qacc + m->nv;
    for (int d = 0; d < m->nv; d++) dd[d] = m->timestep * m->dof_damping[d];
    factor_solve_i(w, w->M, dd, L, qacc, w->efc_Ma);
    advance(w, qacc);
    free(L);
  } else {
    advance(w, w->qacc);
  }
}

/* implicitfast (forward.py:602-610; derivative.py:38-176,178-245,1116-1200): qacc = (M - dt*qDeriv)^-1 Ma with
 * qDeriv = sum_act moment^T (d force / d velocity) moment  -  diag(damping), stateless actuators only */
/* derivative.py:1117-1213 deriv_smooth_vel: Mi = M - dt * qDeriv in the sparsity of M (lower triangle, CSR) */
static void deriv_smooth_vel(W* w, real* Mi) {
  const OrcModel* m = w->m;
  const int nv = m->nv;
  memcpy(Mi, w->M, m->nC * sizeof(real));
  if (m->nu && !(m->disableflags & DSBL_ACTUATION)) {
    for (int a = 0; a < m->nu; a++) {
      real gain = m->actuator_gaintype[a] == GAIN_AFFINE ? m->actuator_gainprm[10 * a + 2] : 0;
      real bias = m->actuator_biastype[a] == BIAS_AFFINE ? m->actuator_biasprm[10 * a + 2] : 0;
      if (bias == 0 && gain == 0) continue;
      if (m->actuator_forcelimited[a]) {
        real f = w->actuator_force[a];
        if (f <= m->actuator_forcerange[2 * a] || f >= m->actuator_forcerange[2 * a + 1]) continue;
      }
      real vel = bias;
      if (gain != 0) { /* derivative.py:142-164: the input the gain multiplies is the activation for a stateful actuator */
        if (m->actuator_dyntype[a] != DYN_NONE) {
          const int last = m->actuator_actadr[a] + m->actuator_actnum[a] - 1;
          vel += gain * (m->actuator_actearly[a] ? next_act(m, a, w->act[last], w->act_dot[last], 1, m->actuator_actlimited[a]) : w->act[last]);
        } else vel += gain * w->ctrl[a];
      }
      if (vel == 0) continue;
      int adr = w->moment_rowadr[a], nnz = w->moment_rownnz[a];
      for (int i = 0; i < nnz; i++) for (int j = 0; j <= i; j++) {
        int di = w->moment_colind[adr + i], dj = w->moment_colind[adr + j];
        /* CSR address of (di, dj): dj must be an ancestor dof of di (M_elemid >= 0) */
        for (int k = 0; k < m->M_rownnz[di]; k++)
          if (m->M_colind[m->M_rowadr[di] + k] == dj) Mi[m->M_rowadr[di] + k] -= m->timestep * w->actuator_moment[adr + i] * w->actuator_moment[adr + j] * vel;
      }
    }
  }
  if (!(m->disableflags & DSBL_DAMPER))
    for (int d = 0; d < nv; d++) Mi[m->M_rowadr[d] + m->M_rownnz[d] - 1] += m->timestep * m->dof_damping[d];
  if (!(m->disableflags & DSBL_DAMPER)) /* derivative.py:262-318: tendon damping, entries of the M sparsity pattern only */
    for (int t = 0; t < m->ntendon; t++) {
      if (m->tendon_damping[t] == 0) continue;
      for (int a = 0; a < m->ten_J_rownnz[t]; a++) for (int b = 0; b < m->ten_J_rownnz[t]; b++) {
        const int di = m->ten_J_colind[m->ten_J_rowadr[t] + a], dj = m->ten_J_colind[m->ten_J_rowadr[t] + b];
        for (int k = 0; k < m->M_rownnz[di]; k++)
          if (m->M_colind[m->M_rowadr[di] + k] == dj) Mi[m->M_rowadr[di] + k] += m->timestep * w->ten_J[m->ten_J_rowadr[t] + a] * w->ten_J[m->ten_J_rowadr[t] + b] * m->tendon_damping[t];
      }
    }
}
static void implicitfast(W* w) {
  const OrcModel* m = w->m;
  const int nv = m->nv;
  int qld = 0; for (int t = 0; t < m->ntree; t++) qld += m->tree_dofnum[t] * m->tree_dofnum[t];
  real* buf = (real*)calloc((size_t)qld + nv + m->nC, sizeof(real));
  real *L = buf, *qacc = buf + qld, *Mi = qacc + nv;
  deriv_smooth_vel(w, Mi);
  factor_solve_i(w, Mi, NULL, L, qacc, w->efc_Ma);
  advance(w, qacc);
  free(buf);
}

/* D-structure (types.py:1343-1347): row i holds the dofs coupled to dof i -- its ancestors, itself, its descendants -- in ascending
 * order; mapM2D sends an entry (i, j) to the entry (max, min) of the lower-triangular M.  Derived here from the M-structure. */
typedef struct { int nD; int *rownnz, *rowadr, *diag, *colind, *mapM2D; } DStruct;
static DStruct dstruct_make(const OrcModel* m) {
  const int nv = m->nv;
  DStruct D;
  int* cnt = (int*)calloc((size_t)nv + 1, sizeof(int));
  for (int i = 0; i < nv; i++)
    for (int k = 0; k < m->M_rownnz[i]; k++) { const int j = m->M_colind[m->M_rowadr[i] + k]; cnt[i]++; if (j != i) cnt[j]++; }
  D.rownnz = (int*)malloc((size_t)(nv + 1) * sizeof(int)); D.rowadr = (int*)malloc((size_t)(nv + 1) * sizeof(int)); D.diag = (int*)malloc((size_t)(nv + 1) * sizeof(int));
  int adr = 0;
  for (int i = 0; i < nv; i++) { D.rownnz[i] = cnt[i]; D.rowadr[i] = adr; adr += cnt[i]; }
  D.nD = adr;
  D.colind = (int*)malloc((size_t)(adr + 1) * sizeof(int)); D.mapM2D = (int*)malloc((size_t)(adr + 1) * sizeof(int));
  memset(cnt, 0, (size_t)(nv + 1) * sizeof(int));
  /* ascending columns: first the row's own M entries (ancestors, then the diagonal), then rows j > i that hold i, in order of j */
  for (int i = 0; i < nv; i++)
    for (int k = 0; k < m->M_rownnz[i]; k++) {
      const int e = m->M_rowadr[i] + k, j = m->M_colind[e];
      if (j == i) D.diag[i] = cnt[i];
      D.colind[D.rowadr[i] + cnt[i]] = j; D.mapM2D[D.rowadr[i] + cnt[i]] = e; cnt[i]++;
    }
  for (int j = 0; j < nv; j++)
    for (int k = 0; k < m->M_rownnz[j] - 1; k++) {
      const int e = m->M_rowadr[j] + k, i = m->M_colind[e];
      D.colind[D.rowadr[i] + cnt[i]] = j; D.mapM2D[D.rowadr[i] + cnt[i]] = e; cnt[i]++;
    }
  free(cnt);
  return D;
}
static void dstruct_free(DStruct* D) { free(D->rownnz); free(D->rowadr); free(D->diag); free(D->colind); free(D->mapM2D); }

/* derivative.py:321-584 deriv_rne_vel: qLU -= dt * d(qfrc_bias) / d(qvel), column by column (dof k), in the D-structure.
 * Forward pass 1 (cvel, cdof_dot), forward pass 2 (cacc, body force), backward accumulation, projection on the joint axes. */
static void deriv_rne_vel_sub(W* w, const DStruct* D, real* qLU) {
  const OrcModel* m = w->m;
  const int nv = m->nv, nb = m->nbody;
  real* Dcvel = (real*)malloc((size_t)6 * (3 * nb + nv) * sizeof(real));
  real *Dcdd = Dcvel + 6 * nb, *Dcacc = Dcdd + 6 * nv, *Dcfrc = Dcacc + 6 * nb;
  for (int k = 0; k < nv; k++) {
    memset(Dcvel, 0, (size_t)6 * (3 * nb + nv) * sizeof(real));
    for (int b = 1; b < nb; b++) {  /* bodies are numbered parents first (derivative.py:336-402) */
      const int pid = m->body_parentid[b];
      real cv[6];
      memcpy(cv, Dcvel + 6 * pid, sizeof cv);
      int dof = m->body_dofadr[b];
      for (int j = m->body_jntadr[b]; j < m->body_jntadr[b] + m->body_jntnum[b]; j++) {
        const int t = m->jnt_type[j];
        if (t == JNT_FREE) {
          if (k >= dof && k < dof + 3) for (int c = 0; c < 6; c++) cv[c] += w->cdof[6 * k + c];
          for (int a = 3; a < 6; a++) motion_cross(cv, w->cdof + 6 * (dof + a), Dcdd + 6 * (dof + a));
          if (k >= dof + 3 && k < dof + 6) for (int c = 0; c < 6; c++) cv[c] += w->cdof[6 * k + c];
          dof += 6;
        } else if (t == JNT_BALL) {
          for (int a = 0; a < 3; a++) motion_cross(cv, w->cdof + 6 * (dof + a), Dcdd + 6 * (dof + a));
          if (k >= dof && k < dof + 3) for (int c = 0; c < 6; c++) cv[c] += w->cdof[6 * k + c];
          dof += 3;
        } else {
          motion_cross(cv, w->cdof + 6 * dof, Dcdd + 6 * dof);
          if (k == dof) for (int c = 0; c < 6; c++) cv[c] += w->cdof[6 * dof + c];
          dof += 1;
        }
      }
      memcpy(Dcvel + 6 * b, cv, sizeof cv);
    }
    for (int b = 1; b < nb; b++) {  /* derivative.py:405-459 */
      const int pid = m->body_parentid[b];
      real dca[6];
      memcpy(dca, Dcacc + 6 * pid, sizeof dca);
      for (int j = m->body_dofadr[b]; j < m->body_dofadr[b] + m->body_dofnum[b]; j++) {
        if (j == k) for (int c = 0; c < 6; c++) dca[c] += w->cdof_dot[6 * j + c];
        for (int c = 0; c < 6; c++) dca[c] += Dcdd[6 * j + c] * w->qvel[j];
      }
      memcpy(Dcacc + 6 * b, dca, sizeof dca);
      real t1[6], icv[6], idcv[6], x1[6], x2[6];
      inert_vec(w->cinert + 10 * b, dca, t1);
      inert_vec(w->cinert + 10 * b, w->cvel + 6 * b, icv);
      inert_vec(w->cinert + 10 * b, Dcvel + 6 * b, idcv);
      motion_cross_force(Dcvel + 6 * b, icv, x1);
      motion_cross_force(w->cvel + 6 * b, idcv, x2);
      for (int c = 0; c < 6; c++) Dcfrc[6 * b + c] = t1[c] + x1[c] + x2[c];
    }
    for (int b = nb - 1; b > 0; b--) {  /* derivative.py:462-479: children into parents */
      const int pid = m->body_parentid[b];
      for (int c = 0; c < 6; c++) Dcfrc[6 * pid + c] += Dcfrc[6 * b + c];
    }
    for (int i = 0; i < nv; i++)  /* derivative.py:482-511: entries (i, k) of the D-structure */
      for (int e = D->rowadr[i]; e < D->rowadr[i] + D->rownnz[i]; e++)
        if (D->colind[e] == k) {
          real s = 0;
          for (int c = 0; c < 6; c++) s += w->cdof[6 * i + c] * Dcfrc[6 * m->dof_bodyid[i] + c];
          qLU[e] -= m->timestep * s;
        }
  }
  free(Dcvel);
}

/* smooth.py:3376-3478 _factor_solve_lu_sparse_fused: in-place sparse LU without fill-in (rows from the last to the first; the part of a
 * row right of the diagonal becomes U with a unit diagonal, the part left of and on the diagonal L), then (U + I) y = b, L x = y */
static void factor_solve_lu(const OrcModel* m, const DStruct* D, real* qLU, real* x, const real* b) {
  const int nv = m->nv;
  int* rem = (int*)malloc((size_t)(nv + 1) * sizeof(int));
  for (int i = 0; i < nv; i++) rem[i] = D->rownnz[i];
  for (int i = nv - 1; i >= 0; i--) {
    const int ii = D->rowadr[i] + rem[i] - 1;
    rem[i]--;
    const real LUii = qLU[ii];
    for (int j = i - 1; j >= 0; j--) {
      const int ji = D->rowadr[j] + rem[j] - 1;
      if (rem[j] > 0 && D->colind[ji] == i) {
        rem[j]--;
        const real LUji = qLU[ji] / LUii;
        qLU[ji] = LUji;
        int ic = D->rowadr[i], jc = D->rowadr[j];
        const int jend = D->rowadr[j] + rem[j], iend = D->rowadr[i] + D->rownnz[i];
        while (jc < jend && ic < iend) {
          const int ci = D->colind[ic], cj = D->colind[jc];
          if (ci == cj) { qLU[jc] -= qLU[ic] * LUji; ic++; jc++; }
          else if (ci > cj) jc++;
          else ic++;
        }
      }
    }
  }
  for (int i = nv - 1; i >= 0; i--) {
    real acc = b[i];
    for (int e = D->rowadr[i] + D->diag[i] + 1; e < D->rowadr[i] + D->rownnz[i]; e++) acc -= qLU[e] * x[D->colind[e]];
    x[i] = acc;
  }
  for (int i = 0; i < nv; i++) {
    real acc = x[i];
    for (int e = D->rowadr[i]; e < D->rowadr[i] + D->diag[i]; e++) acc -= qLU[e] * x[D->colind[e]];
    x[i] = acc / qLU[D->rowadr[i] + D->diag[i]];
  }
  free(rem);
}

/* forward.py:578-600 implicit, IMPLICIT branch: qLU = M - dt (qDeriv_smooth + d RNE / d qvel) in the D-structure, qacc = qLU \ Ma */
static void implicit_full(W* w) {
  const OrcModel* m = w->m;
  const int nv = m->nv;
  DStruct D = dstruct_make(m);
  real* Mi = (real*)calloc((size_t)m->nC + D.nD + nv + 1, sizeof(real));
  real *qLU = Mi + m->nC, *qacc = qLU + D.nD;
  deriv_smooth_vel(w, Mi);
  for (int e = 0; e < D.nD; e++) qLU[e] = Mi[D.mapM2D[e]];
  deriv_rne_vel_sub(w, &D, qLU);
  factor_solve_lu(m, &D, qLU, qacc, w->efc_Ma);
  advance(w, qacc);
  free(Mi);
  dstruct_free(&D);
}

/* ------------------------------------------------------------------ sensors (sensor.py; smooth.py:3500-3612 subtree_vel, :1743 rne_postconstraint) */
enum { SENS_TOUCH = 0, SENS_ACCELEROMETER, SENS_VELOCIMETER, SENS_GYRO, SENS_FORCE, SENS_TORQUE, SENS_MAGNETOMETER, SENS_RANGEFINDER, SENS_CAMPROJECTION,
  SENS_JOINTPOS, SENS_JOINTVEL, SENS_TENDONPOS, SENS_TENDONVEL, SENS_ACTUATORPOS, SENS_ACTUATORVEL, SENS_ACTUATORFRC, SENS_JOINTACTFRC, SENS_TENDONACTFRC,
  SENS_BALLQUAT, SENS_BALLANGVEL, SENS_JOINTLIMITPOS, SENS_JOINTLIMITVEL, SENS_JOINTLIMITFRC, SENS_TENDONLIMITPOS, SENS_TENDONLIMITVEL, SENS_TENDONLIMITFRC,
  SENS_FRAMEPOS, SENS_FRAMEQUAT, SENS_FRAMEXAXIS, SENS_FRAMEYAXIS, SENS_FRAMEZAXIS, SENS_FRAMELINVEL, SENS_FRAMEANGVEL, SENS_FRAMELINACC, SENS_FRAMEANGACC,
  SENS_SUBTREECOM, SENS_SUBTREELINVEL, SENS_SUBTREEANGMOM, SENS_INSIDESITE, SENS_GEOMDIST, SENS_GEOMNORMAL, SENS_GEOMFROMTO, SENS_CONTACT, SENS_E_POTENTIAL,
  SENS_E_KINETIC, SENS_CLOCK };
enum { STAGE_POS = 1, STAGE_VEL = 2, STAGE_ACC = 3 };
enum { DATATYPE_REAL = 0, DATATYPE_POSITIVE = 1 };
enum { OBJ_BODY = 1, OBJ_XBODY = 2, OBJ_GEOM = 5, OBJ_SITE = 6, OBJ_CAMERA = 7 };
#define DSBL_SENSOR (1 << 13)

/* sensor.py:56-113 _write_scalar / _write_vector: cutoff clamps REAL data to [-c, c] and POSITIVE data to (-inf, c] */
static void sensor_write(W* w, int s, int dim, const real* v) {
  const OrcModel* m = w->m;
  real cutoff = m->sensor_cutoff[s];
  for (int i = 0; i < dim; i++) {
    real x = v[i];
    if (cutoff > 0) {
      if (m->sensor_datatype[s] == DATATYPE_REAL) x = rclamp(x, -cutoff, cutoff);
      else if (m->sensor_datatype[s] == DATATYPE_POSITIVE) x = rmin(x, cutoff);
    }
    w->sensordata[m->sensor_adr[s] + i] = x;
  }
}
/* smooth.py:3500-3612: subtree linear velocity and angular momentum about the subtree centre of mass */
static void subtree_vel(W* w) {
  const OrcModel* m = w->m;
  const int nb = m->nbody;
  real* bodyvel = (real*)malloc((size_t)6 * nb * sizeof(real));
  for (int b = 0; b < nb; b++) {
    const real *cv = w->cvel + 6 * b, *ximat = w->ximat + 9 * b;
    real dif[3], cr[3], lin[3], dv[3];
    v3sub(w->xipos + 3 * b, w->subtree_com + 3 * m->body_rootid[b], dif);
    cross3(dif, cv, cr);
    for (int i = 0; i < 3; i++) lin[i] = cv[3 + i] - cr[i];
    for (int i = 0; i < 3; i++) w->subtree_linvel[3 * b + i] = m->body_mass[b] * lin[i];
    matT_vec3(ximat, cv, dv);
    for (int i = 0; i < 3; i++) dv[i] *= m->body_inertia[3 * b + i];
    matvec3(ximat, dv, w->subtree_angmom + 3 * b);
    for (int i = 0; i < 3; i++) { bodyvel[6 * b + i] = cv[i]; bodyvel[6 * b + 3 + i] = lin[i]; }
  }
  for (int b = nb - 1; b >= 0; b--) { /* _linear_momentum, deepest bodies first */
    if (b) for (int i = 0; i < 3; i++) w->subtree_linvel[3 * m->body_parentid[b] + i] += w->subtree_linvel[3 * b + i];
    for (int i = 0; i < 3; i++) w->subtree_linvel[3 * b + i] /= rmax(MJ_MINVAL, m->body_subtreemass[b]);
  }
  for (int b = nb - 1; b >= 1; b--) { /* _angular_momentum */
    const int pid = m->body_parentid[b];
    real dx[3], dv[3], dL[3];
    v3sub(w->xipos + 3 * b, w->subtree_com + 3 * b, dx);
    for (int i = 0; i < 3; i++) dv[i] = (bodyvel[6 * b + 3 + i] - w->subtree_linvel[3 * b + i]) * m->body_mass[b];
    cross3(dx, dv, dL);
    for (int i = 0; i < 3; i++) w->subtree_angmom[3 * b + i] += dL[i];
    for (int i = 0; i < 3; i++) w->subtree_angmom[3 * pid + i] += w->subtree_angmom[3 * b + i];
    v3sub(w->subtree_com + 3 * b, w->subtree_com + 3 * pid, dx);
    for (int i = 0; i < 3; i++) dv[i] = (w->subtree_linvel[3 * b + i] - w->subtree_linvel[3 * pid + i]) * m->body_subtreemass[b];
    cross3(dx, dv, dL);
    for (int i = 0; i < 3; i++) w->subtree_angmom[3 * pid + i] += dL[i];
  }
  free(bodyvel);
}
/* support.py:326-397 contact_force_fn with to_world_frame: 6D (force, torque) of one contact in the world frame */
static void contact_wrench_world(const W* w, int c, real* out) {
  const OrcModel* m = w->m;
  real f[6] = {0, 0, 0, 0, 0, 0};
  const int dim = w->con_dim[c], *adr = w->con_efc_address + m->nmaxpyramid * c;
  if (adr[0] >= 0) {
    if (m->cone == CONE_PYRAMIDAL) {
      if (dim == 1) f[0] = w->efc_force[adr[0]];
      else for (int i = 0; i < dim - 1; i++) {
        int a = 2 * i + adr[0];
        real d1 = a < w->njmax ? w->efc_force[a] : 0, d2 = a + 1 < w->njmax ? w->efc_force[a + 1] : 0;
        f[0] += d1 + d2; f[i + 1] = (d1 - d2) * w->con_friction[5 * c + i];
      }
    } else for (int i = 0; i < dim; i++) if (adr[i] < w->njmax) f[i] = w->efc_force[adr[i]];
  }
  const real* R = w->con_frame + 9 * c;
  for (int k = 0; k < 3; k++) { out[k] = f[0] * R[k] + f[1] * R[3 + k] + f[2] * R[6 + k]; out[3 + k] = f[3] * R[k] + f[4] * R[3 + k] + f[5] * R[6 + k]; }
}
/* support.py:476-485 transform_force: (torque - offset x force, force) as a spatial (angular, linear) vector */
static void transform_force(const real* force, const real* torque, const real* offset, real* o) {
  real cr[3]; cross3(offset, force, cr);
  for (int i = 0; i < 3; i++) { o[i] = torque[i] - cr[i]; o[3 + i] = force[i]; }
}
/* smooth.py:1743 rne_postconstraint: cfrc_ext from applied wrenches (:1518), connect / weld equalities between bodies (:1562) and
 * contacts (:1660); cacc including qacc (:1364-1425 with flg_acc); cfrc_int (:1428) accumulated up the tree (:1453) */
static void rne_postconstraint(W* w) {
  const OrcModel* m = w->m;
  const int nb = m->nbody;
  memset(w->cfrc_ext, 0, 6 * sizeof(real));
  for (int b = 1; b < nb; b++) {
    real off[3]; v3sub(w->subtree_com + 3 * m->body_rootid[b], w->xipos + 3 * b, off);
    transform_force(w->xfrc_applied + 6 * b, w->xfrc_applied + 6 * b + 3, off, w->cfrc_ext + 6 * b);
  }
  for (int e = 0; e < w->ne[0];) { /* rows are ordered connect (3 rows each), weld (6), joint (1): constraint.py:4909-5010 */
    const int id = w->efc_id[e], type = m->eq_type[id];
    if (type != EQ_CONNECT && type != EQ_WELD) break;
    const int nrow = type == EQ_CONNECT ? 3 : 6, b1 = m->eq_obj1id[id], b2 = m->eq_obj2id[id];
    real force[3] = {w->efc_force[e], w->efc_force[e + 1], w->efc_force[e + 2]}, torque[3] = {0, 0, 0}, c6[6], pos[3], dif[3], t[3];
    if (type == EQ_WELD) for (int i = 0; i < 3; i++) torque[i] = w->efc_force[e + 3 + i];
    const real* data = m->eq_data + 11 * id;
    if (b1) {
      matvec3(w->xmat + 9 * b1, type == EQ_CONNECT ? data : data + 3, t);
      for (int i = 0; i < 3; i++) pos[i] = t[i] + w->xpos[3 * b1 + i];
      v3sub(w->subtree_com + 3 * m->body_rootid[b1], pos, dif);
      transform_force(force, torque, dif, c6);
      for (int i = 0; i < 6; i++) w->cfrc_ext[6 * b1 + i] += c6[i];
    }
    if (b2) {
      matvec3(w->xmat + 9 * b2, type == EQ_CONNECT ? data + 3 : data, t);
      for (int i = 0; i < 3; i++) pos[i] = t[i] + w->xpos[3 * b2 + i];
      v3sub(w->subtree_com + 3 * m->body_rootid[b2], pos, dif);
      transform_force(force, torque, dif, c6);
      for (int i = 0; i < 6; i++) w->cfrc_ext[6 * b2 + i] -= c6[i];
    }
    e += nrow;
  }
  const int ncon = w->ncon[0] < w->nconmax ? w->ncon[0] : w->nconmax;
  for (int c = 0; c < ncon; c++) {
    const int id1 = m->geom_bodyid[w->con_geom[2 * c]], id2 = m->geom_bodyid[w->con_geom[2 * c + 1]];
    if (id1 == 0 && id2 == 0) continue;
    real wr[6], off[3], c6[6];
    contact_wrench_world(w, c, wr);
    if (id1) { v3sub(w->subtree_com + 3 * m->body_rootid[id1], w->con_pos + 3 * c, off); transform_force(wr, wr + 3, off, c6); for (int i = 0; i < 6; i++) w->cfrc_ext[6 * id1 + i] -= c6[i]; }
    if (id2) { v3sub(w->subtree_com + 3 * m->body_rootid[id2], w->con_pos + 3 * c, off); transform_force(wr, wr + 3, off, c6); for (int i = 0; i < 6; i++) w->cfrc_ext[6 * id2 + i] += c6[i]; }
  }
  memset(w->cacc, 0, 6 * sizeof(real));
  if (!(m->disableflags & DSBL_GRAVITY)) for (int i = 0; i < 3; i++) w->cacc[3 + i] = -m->gravity[i];
  for (int b = 1; b < nb; b++) {
    real a[6];
    memcpy(a, w->cacc + 6 * m->body_parentid[b], sizeof a);
    for (int k = 0; k < m->body_dofnum[b]; k++) {
      int d = m->body_dofadr[b] + k;
      for (int i = 0; i < 6; i++) a[i] += w->cdof_dot[6 * d + i] * w->qvel[d];
      for (int i = 0; i < 6; i++) a[i] += w->cdof[6 * d + i] * w->qacc[d];
    }
    memcpy(w->cacc + 6 * b, a, sizeof a);
  }
  memset(w->cfrc_int, 0, 6 * sizeof(real));
  for (int b = 1; b < nb; b++) {
    real f[6], iv[6], g[6];
    inert_vec(w->cinert + 10 * b, w->cacc + 6 * b, f);
    inert_vec(w->cinert + 10 * b, w->cvel + 6 * b, iv);
    motion_cross_force(w->cvel + 6 * b, iv, g);
    for (int i = 0; i < 6; i++) w->cfrc_int[6 * b + i] = f[i] + g[i] - w->cfrc_ext[6 * b + i];
  }
  for (int b = nb - 1; b >= 1; b--) { int p = m->body_parentid[b]; for (int i = 0; i < 6; i++) w->cfrc_int[6 * p + i] += w->cfrc_int[6 * b + i]; }
}
/* object frame position / orientation (sensor.py:266-317 _get_pos / _get_mat) */
static const real* obj_pos(const W* w, int objtype, int id) {
  switch (objtype) {
    case OBJ_BODY: return w->xipos + 3 * id;
    case OBJ_XBODY: return w->xpos + 3 * id;
    case OBJ_GEOM: return w->geom_xpos + 3 * id;
    case OBJ_SITE: return w->site_xpos + 3 * id;
    default: return w->cam_xpos + 3 * id;
  }
}
static const real* obj_mat(const W* w, int objtype, int id) {
  switch (objtype) {
    case OBJ_BODY: return w->ximat + 9 * id;
    case OBJ_XBODY: return w->xmat + 9 * id;
    case OBJ_GEOM: return w->geom_xmat + 9 * id;
    case OBJ_SITE: return w->site_xmat + 9 * id;
    default: return w->cam_xmat + 9 * id;
  }
}
/* ray.py:106 _ray_quad: smallest non-negative root of a x^2 + 2 b x + c = 0 (and both roots), -1 if none */
static real ray_quad(real a, real b, real c, real* x2) {
  real det = b * b - a * c;
  x2[0] = x2[1] = -1;
  if (det < MJ_MINVAL) return -1;
  det = (real)sqrt((double)det);
  real den = a != 0 ? 1 / a : 0; /* safe_div */
  x2[0] = (-b - det) * den; x2[1] = (-b + det) * den;
  return x2[0] >= 0 ? x2[0] : (x2[1] >= 0 ? x2[1] : -1);
}
static real ray_sphere(const real* pos, real dist_sqr, const real* pnt, const real* vec) { /* ray.py:238 */
  real dif[3], xx[2]; v3sub(pnt, pos, dif);
  return ray_quad(dot3(vec, vec), dot3(vec, dif), dot3(dif, dif) - dist_sqr, xx);
}
/* ray.py:799 ray_geom, distance only, for the shapes a site can have (sphere, capsule, ellipsoid, cylinder, box) */
static real ray_geom_dist(const real* pos, const real* mat, const real* size, const real* pnt, const real* vec, int type) {
  real xx[2];
  if (type == GEOM_SPHERE) return ray_sphere(pos, size[0] * size[0], pnt, vec);
  real d[3], lpnt[3], lvec[3];
  v3sub(pnt, pos, d); matT_vec3(mat, d, lpnt); matT_vec3(mat, vec, lvec); /* :33 _ray_map */
  if (type == GEOM_CAPSULE) { /* :255 */
    real ssz = size[0] + size[1];
    if (ray_sphere(pos, ssz * ssz, pnt, vec) < 0) return -1;
    real x = -1, sq = size[0] * size[0];
    real a = lvec[0] * lvec[0] + lvec[1] * lvec[1], b = lvec[0] * lpnt[0] + lvec[1] * lpnt[1], c = lpnt[0] * lpnt[0] + lpnt[1] * lpnt[1] - sq;
    real sol = ray_quad(a, b, c, xx);
    if (sol >= 0 && rabs(lpnt[2] + sol * lvec[2]) <= size[1]) if (x < 0 || sol < x) x = sol;
    real ldif[3] = {lpnt[0], lpnt[1], lpnt[2] - size[1]};
    a += lvec[2] * lvec[2]; b = dot3(lvec, ldif); c = dot3(ldif, ldif) - sq;
    ray_quad(a, b, c, xx);
    for (int i = 0; i < 2; i++) if (xx[i] >= 0 && lpnt[2] + xx[i] * lvec[2] >= size[1]) if (x < 0 || xx[i] < x) x = xx[i];
    ldif[2] = lpnt[2] + size[1];
    b = dot3(lvec, ldif); c = dot3(ldif, ldif) - sq;
    ray_quad(a, b, c, xx);
    for (int i = 0; i < 2; i++) if (xx[i] >= 0 && lpnt[2] + xx[i] * lvec[2] <= -size[1]) if (x < 0 || xx[i] < x) x = xx[i];
    return x;
  }
  if (type == GEOM_ELLIPSOID) { /* :329 */
    real sv[3], sp[3], a = 0, b = 0, c = -1;
    for (int i = 0; i < 3; i++) { real q = size[i] * size[i], si = q != 0 ? 1 / q : 0; sv[i] = si * lvec[i]; sp[i] = si * lpnt[i]; }
    a = dot3(sv, lvec); b = dot3(sv, lpnt); c = dot3(sp, lpnt) - 1;
    return ray_quad(a, b, c, xx);
  }
  if (type == GEOM_CYLINDER) { /* :360 */
    if (ray_sphere(pos, size[0] * size[0] + size[1] * size[1], pnt, vec) < 0) return -1;
    real x = -1;
    if (rabs(lvec[2]) > MJ_MINVAL)
      for (int side = -1; side <= 1; side += 2) {
        real sol = ((real)side * size[1] - lpnt[2]) / lvec[2];
        if (sol >= 0) {
          real p0 = lpnt[0] + sol * lvec[0], p1 = lpnt[1] + sol * lvec[1];
          if (p0 * p0 + p1 * p1 <= size[0] * size[0]) if (x < 0 || sol < x) x = sol;
        }
      }
    real a = lvec[0] * lvec[0] + lvec[1] * lvec[1], b = lvec[0] * lpnt[0] + lvec[1] * lpnt[1], c = lpnt[0] * lpnt[0] + lpnt[1] * lpnt[1] - size[0] * size[0];
    real sol = ray_quad(a, b, c, xx);
    if (sol >= 0 && rabs(lpnt[2] + sol * lvec[2]) <= size[1]) if (x < 0 || sol < x) x = sol;
    return x;
  }
  if (type == GEOM_BOX) { /* :421 */
    if (ray_sphere(pos, dot3(size, size), pnt, vec) < 0) return -1;
    real x = -1;
    for (int i = 0; i < 3; i++) {
      if (rabs(lvec[i]) <= MJ_MINVAL) continue;
      for (int side = -1; side <= 1; side += 2) {
        real sol = ((real)side * size[i] - lpnt[i]) / lvec[i];
        if (sol < 0) continue;
        int id0 = i == 0 ? 1 : 0, id1 = i == 2 ? 1 : 2;
        real p0 = lpnt[id0] + sol * lvec[id0], p1 = lpnt[id1] + sol * lvec[id1];
        if (rabs(p0) <= size[id0] && rabs(p1) <= size[id1]) if (x < 0 || sol < x) x = sol;
      }
    }
    return x;
  }
  return -1;
}
/* sensor.py:2063 _sensor_touch: sum of the normal forces of the sensorised body's contacts whose force ray meets the site volume */
static real sensor_touch(const W* w, int site) {
  const OrcModel* m = w->m;
  const int body = m->site_bodyid[site], ncon = w->ncon[0] < w->nconmax ? w->ncon[0] : w->nconmax;
  real total = 0;
  for (int c = 0; c < ncon; c++) {
    const int b1 = m->geom_bodyid[w->con_geom[2 * c]], b2 = m->geom_bodyid[w->con_geom[2 * c + 1]];
    const int* adr = w->con_efc_address + m->nmaxpyramid * c;
    if (adr[0] < 0 || (body != b1 && body != b2)) continue;
    real nf = w->efc_force[adr[0]];
    if (m->cone == CONE_PYRAMIDAL) for (int i = 1; i < 2 * (w->con_dim[c] - 1); i++) nf += w->efc_force[adr[i]];
    if (nf <= 0) continue;
    real ray[3] = {w->con_frame[9 * c] * nf, w->con_frame[9 * c + 1] * nf, w->con_frame[9 * c + 2] * nf};
    normalize3(ray);
    if (body == b2) for (int i = 0; i < 3; i++) ray[i] = -ray[i];
    if (ray_geom_dist(w->site_xpos + 3 * site, w->site_xmat + 9 * site, m->site_size + 3 * site, w->con_pos + 3 * c, ray, m->site_type[site]) >= 0) total += nf;
  }
  return total;
}
/* body that carries an object (sensor.py:1066-1105 _cvel_offset / :320 _get_body_id) */
static int obj_body(const OrcModel* m, int objtype, int id) {
  switch (objtype) {
    case OBJ_BODY: case OBJ_XBODY: return id;
    case OBJ_GEOM: return m->geom_bodyid[id];
    case OBJ_SITE: return m->site_bodyid[id];
    case OBJ_CAMERA: return m->cam_bodyid[id];
    default: return 0;
  }
}
/* sensor.py:810 sensor_pos, :1432 sensor_vel, :2512 sensor_acc for the sensor types this build carries */
static void sensors(W* w, int stage) {
  const OrcModel* m = w->m;
  if (!m->nsensor || (m->disableflags & DSBL_SENSOR)) return;
  int need_subtree = 0, need_cacc = 0;
  for (int s = 0; s < m->nsensor; s++) {
    int t = m->sensor_type[s];
    if (t == SENS_SUBTREELINVEL || t == SENS_SUBTREEANGMOM) need_subtree = 1;
    if (t == SENS_ACCELEROMETER || t == SENS_FORCE || t == SENS_TORQUE || t == SENS_FRAMELINACC || t == SENS_FRAMEANGACC) need_cacc = 1;
  }
  if (stage == STAGE_VEL && need_subtree) subtree_vel(w);
  if (stage == STAGE_ACC && need_cacc) rne_postconstraint(w);
  for (int s = 0; s < m->nsensor; s++) {
    if (m->sensor_needstage[s] != stage) continue;
    const int t = m->sensor_type[s], id = m->sensor_objid[s], rt = m->sensor_reftype[s], rid = m->sensor_refid[s];
    real v[4] = {0, 0, 0, 0};
    switch (t) {
      case SENS_JOINTPOS: v[0] = w->qpos[m->jnt_qposadr[id]]; break;
      case SENS_TENDONPOS: v[0] = w->ten_length[id]; break;
      case SENS_ACTUATORPOS: v[0] = w->actuator_length[id]; break;
      case SENS_BALLQUAT: { memcpy(v, w->qpos + m->jnt_qposadr[id], 4 * sizeof(real)); normalize4(v); break; }
      case SENS_FRAMEPOS: { /* sensor.py:377: relative to the reference frame when one is given */
        memcpy(v, obj_pos(w, m->sensor_objtype[s], id), 3 * sizeof(real));
        if (rid > -1) { real dlt[3]; v3sub(v, obj_pos(w, rt, rid), dlt); matT_vec3(obj_mat(w, rt, rid), dlt, v); }
        break; }
      case SENS_FRAMEXAXIS: case SENS_FRAMEYAXIS: case SENS_FRAMEZAXIS: {
        const real* R = obj_mat(w, m->sensor_objtype[s], id); int c = t - SENS_FRAMEXAXIS;
        v[0] = R[c]; v[1] = R[3 + c]; v[2] = R[6 + c];
        if (rid > -1) { real ax[3] = {v[0], v[1], v[2]}; matT_vec3(obj_mat(w, rt, rid), ax, v); } /* sensor.py:406 */
        break; }
      case SENS_FRAMEQUAT: { /* sensor.py:342-374 _get_quat */
        const int ot = m->sensor_objtype[s];
        const real* local = ot == OBJ_BODY ? m->body_iquat + 4 * id : ot == OBJ_GEOM ? m->geom_quat + 4 * id : ot == OBJ_SITE ? m->site_quat + 4 * id : ot == OBJ_CAMERA ? m->cam_quat + 4 * id : NULL;
        const real* xq = w->xquat + 4 * obj_body(m, ot, id);
        if (local) mul_quat(xq, local, v); else memcpy(v, xq, 4 * sizeof(real));
        if (rid > -1) { /* sensor.py:470-482: conj(refquat) * quat */
          const real* rl = rt == OBJ_BODY ? m->body_iquat + 4 * rid : rt == OBJ_GEOM ? m->geom_quat + 4 * rid : rt == OBJ_SITE ? m->site_quat + 4 * rid : rt == OBJ_CAMERA ? m->cam_quat + 4 * rid : NULL;
          const real* rx = w->xquat + 4 * obj_body(m, rt, rid);
          real rq[4], q[4] = {v[0], v[1], v[2], v[3]};
          if (rl) mul_quat(rx, rl, rq); else memcpy(rq, rx, sizeof rq);
          rq[1] = -rq[1]; rq[2] = -rq[2]; rq[3] = -rq[3];
          mul_quat(rq, q, v);
        }
        break; }
      case SENS_SUBTREECOM: memcpy(v, w->subtree_com + 3 * id, 3 * sizeof(real)); break;
      case SENS_CLOCK: v[0] = w->time[0]; break;
      case SENS_JOINTLIMITPOS: case SENS_JOINTLIMITVEL: case SENS_JOINTLIMITFRC: { /* sensor.py:228, :1028, :1640: the joint's active limit row, else 0 */
        for (int e = w->ne[0] + w->nf[0]; e < w->ne[0] + w->nf[0] + w->nl[0] && e < w->njmax; e++)
          if (w->efc_id[e] == id && w->efc_type[e] == CNSTR_LIMIT_JOINT)
            v[0] = t == SENS_JOINTLIMITPOS ? w->efc_pos[e] - w->efc_margin[e] : (t == SENS_JOINTLIMITVEL ? w->efc_vel[e] : w->efc_force[e]);
        break; }
      case SENS_JOINTVEL: v[0] = w->qvel[m->jnt_dofadr[id]]; break;
      case SENS_TENDONVEL: v[0] = w->ten_velocity[id]; break;
      case SENS_ACTUATORVEL: v[0] = w->actuator_velocity[id]; break;
      case SENS_BALLANGVEL: memcpy(v, w->qvel + m->jnt_dofadr[id], 3 * sizeof(real)); break;
      case SENS_FRAMELINVEL: case SENS_FRAMEANGVEL: { /* sensor.py:1108-1293 without a reference frame */
        const int ot = m->sensor_objtype[s], b = obj_body(m, ot, id); const real* cv = w->cvel + 6 * b;
        real off[3], cr[3], lin[3];
        v3sub(obj_pos(w, ot, id), w->subtree_com + 3 * m->body_rootid[b], off); cross3(off, cv, cr);
        for (int i = 0; i < 3; i++) lin[i] = cv[3 + i] - cr[i];
        if (rid > -1) { /* sensor.py:1188-1210, :1255-1291: velocity relative to, and expressed in, the reference frame */
          const int rb = obj_body(m, rt, rid); const real* rv = w->cvel + 6 * rb;
          real roff[3], rlin[3], rvec[3], cr2[3], rel[3];
          v3sub(obj_pos(w, rt, rid), w->subtree_com + 3 * m->body_rootid[rb], roff); cross3(roff, rv, cr);
          for (int i = 0; i < 3; i++) rlin[i] = rv[3 + i] - cr[i];
          v3sub(obj_pos(w, ot, id), obj_pos(w, rt, rid), rvec); cross3(rvec, rv, cr2);
          for (int i = 0; i < 3; i++) rel[i] = t == SENS_FRAMEANGVEL ? cv[i] - rv[i] : lin[i] - rlin[i] + cr2[i];
          matT_vec3(obj_mat(w, rt, rid), rel, v);
        } else memcpy(v, t == SENS_FRAMEANGVEL ? cv : lin, 3 * sizeof(real));
        break; }
      case SENS_SUBTREELINVEL: memcpy(v, w->subtree_linvel + 3 * id, 3 * sizeof(real)); break;
      case SENS_SUBTREEANGMOM: memcpy(v, w->subtree_angmom + 3 * id, 3 * sizeof(real)); break;
      case SENS_GYRO: matT_vec3(w->site_xmat + 9 * id, w->cvel + 6 * m->site_bodyid[id], v); break; /* sensor.py:989 */
      case SENS_VELOCIMETER: { /* sensor.py:964 */
        const int b = m->site_bodyid[id]; const real* cv = w->cvel + 6 * b;
        real dif[3], cr[3], lin[3];
        v3sub(w->site_xpos + 3 * id, w->subtree_com + 3 * m->body_rootid[b], dif); cross3(dif, cv, cr);
        for (int i = 0; i < 3; i++) lin[i] = cv[3 + i] - cr[i];
        matT_vec3(w->site_xmat + 9 * id, lin, v); break; }
      case SENS_ACCELEROMETER: { /* sensor.py:1510 */
        const int b = m->site_bodyid[id]; const real *cv = w->cvel + 6 * b, *ca = w->cacc + 6 * b, *R = w->site_xmat + 9 * id;
        real dif[3], cr[3], t1[3], ang[3], lin[3], acc[3], corr[3];
        v3sub(w->site_xpos + 3 * id, w->subtree_com + 3 * m->body_rootid[b], dif);
        matT_vec3(R, cv, ang);
        cross3(dif, cv, cr); for (int i = 0; i < 3; i++) t1[i] = cv[3 + i] - cr[i]; matT_vec3(R, t1, lin);
        cross3(dif, ca, cr); for (int i = 0; i < 3; i++) t1[i] = ca[3 + i] - cr[i]; matT_vec3(R, t1, acc);
        cross3(ang, lin, corr);
        for (int i = 0; i < 3; i++) v[i] = acc[i] + corr[i]; break; }
      case SENS_FRAMELINACC: case SENS_FRAMEANGACC: { /* sensor.py:1678-1753 */
        const int ot = m->sensor_objtype[s], b = obj_body(m, ot, id); const real *cv = w->cvel + 6 * b, *ca = w->cacc + 6 * b;
        if (t == SENS_FRAMEANGACC) { memcpy(v, ca, 3 * sizeof(real)); break; }
        real off[3], cr[3], lin[3], acc[3], corr[3];
        v3sub(obj_pos(w, ot, id), w->subtree_com + 3 * m->body_rootid[b], off);
        cross3(off, cv, cr); for (int i = 0; i < 3; i++) lin[i] = cv[3 + i] - cr[i];
        cross3(off, ca, cr); for (int i = 0; i < 3; i++) acc[i] = ca[3 + i] - cr[i];
        cross3(cv, lin, corr);
        for (int i = 0; i < 3; i++) v[i] = acc[i] + corr[i];
        break; }
      case SENS_TOUCH: v[0] = sensor_touch(w, id); break;
      case SENS_FORCE: matT_vec3(w->site_xmat + 9 * id, w->cfrc_int + 6 * m->site_bodyid[id] + 3, v); break; /* sensor.py:1542 */
      case SENS_TORQUE: { /* sensor.py:1559 */
        const int b = m->site_bodyid[id]; const real* cf = w->cfrc_int + 6 * b;
        real dif[3], cr[3], t1[3];
        v3sub(w->site_xpos + 3 * id, w->subtree_com + 3 * m->body_rootid[b], dif); cross3(dif, cf + 3, cr);
        for (int i = 0; i < 3; i++) t1[i] = cf[i] - cr[i];
        matT_vec3(w->site_xmat + 9 * id, t1, v); break; }
      case SENS_ACTUATORFRC: v[0] = w->actuator_force[id]; break;
      case SENS_JOINTACTFRC: v[0] = w->qfrc_actuator[m->jnt_dofadr[id]]; break;
      default: continue; /* put_model rejects other types */
    }
    sensor_write(w, s, m->sensor_dim[s], v);
  }
}

static void forward_world(W* w) {
  kinematics(w); com_pos(w); camlight(w); tendon(w); crb(w);
  collision(w); make_constraint(w); transmission(w);
  sensors(w, STAGE_POS);
  fwd_velocity(w); sensors(w, STAGE_VEL); fwd_actuation(w); fwd_acceleration(w);
  solve(w);
  sensors(w, STAGE_ACC);
}

/* forward.py:523-555 rungekutta4 (stateless actuators): called after the first forward() of the step */
static void rungekutta4(W* w) {
  const OrcModel* m = w->m;
  const int nq = m->nq, nv = m->nv;
  const real A[3] = {(real)0.5, (real)0.5, (real)1.0}, B[4] = {(real)(1.0 / 6.0), (real)(1.0 / 3.0), (real)(1.0 / 3.0), (real)(1.0 / 6.0)};
  const int na = m->na;
  real* buf = (real*)calloc((size_t)2 * nq + 3 * nv + 2 * na, sizeof(real));
  real *qpos_t0 = buf, *qpos_new = buf + nq, *qvel_t0 = buf + 2 * nq, *qvel_rk = qvel_t0 + nv, *qacc_rk = qvel_rk + nv, *act_t0 = qacc_rk + nv, *act_dot_rk = act_t0 + na;
  memcpy(qpos_t0, w->qpos, nq * sizeof(real)); memcpy(qvel_t0, w->qvel, nv * sizeof(real)); memcpy(act_t0, w->act, na * sizeof(real));
  for (int d = 0; d < nv; d++) { qvel_rk[d] += B[0] * w->qvel[d]; qacc_rk[d] += B[0] * w->qacc[d]; }
  for (int j = 0; j < na; j++) act_dot_rk[j] += B[0] * w->act_dot[j];
  for (int i = 0; i < 3; i++) {
    /* _rk_perturb_state: the position step uses the current stage velocity, then the velocity is perturbed */
    next_position(m, qpos_t0, w->qvel, A[i], qpos_new);
    memcpy(w->qpos, qpos_new, nq * sizeof(real));
    for (int d = 0; d < nv; d++) w->qvel[d] = qvel_t0[d] + A[i] * w->qacc[d] * m->timestep;
    if (na) next_activation(w, act_t0, A[i], 0); /* forward.py:445-463: unclamped stage activations */
    forward_world(w);
    for (int d = 0; d < nv; d++) { qvel_rk[d] += B[i + 1] * w->qvel[d]; qacc_rk[d] += B[i + 1] * w->qacc[d]; }
    for (int j = 0; j < na; j++) act_dot_rk[j] += B[i + 1] * w->act_dot[j];
  }
  memcpy(w->qpos, qpos_t0, nq * sizeof(real)); memcpy(w->qvel, qvel_t0, nv * sizeof(real));
  if (na) { memcpy(w->act, act_t0, na * sizeof(real)); memcpy(w->act_dot, act_dot_rk, na * sizeof(real)); } /* forward.py:553-555 */
  advance2(w, qacc_rk, qvel_rk);
  free(buf);
}

static int run(const OrcModel* m, OrcData* d, int nthreads, int do_step) {
  if (check_fields(m, d)) return -1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
  for (int wi = 0; wi < d->nworld; wi++) {
    W w;
    make_view(m, d, wi, &w);
    forward_world(&w);
    if (do_step) { if (m->integrator == INT_IMPLICITFAST) implicitfast(&w); else if (m->integrator == INT_IMPLICIT) implicit_full(&w); else if (m->integrator == INT_RK4) rungekutta4(&w); else euler(&w); }
  }
  return 0;
}
int orc_forward(const OrcModel* m, OrcData* d, int nthreads) { return run(m, d, nthreads, 0); }
int orc_step(const OrcModel* m, OrcData* d, int nthreads) { return run(m, d, nthreads, 1); }
