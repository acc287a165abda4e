/* oracle.h -- TEST INFRASTRUCTURE ONLY (not the product path).
 *
 * CPU restatement (plain C, one world at a time, OpenMP over worlds) of the batched physics step
 * of google-deepmind/mujoco_warp: forward.py:1341-1380 (forward/step), smooth.py, collision_driver.py,
 * collision_primitive*.py, constraint.py, solver.py.  Every function in oracle.c cites the reference
 * file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * PARITY PINNING: pinned against outputs of the reference itself.  The reference's unmodified
 * Python sources are executed on the CPU through tools/warp_shim.py (a pure-Python stand-in for the warp API), which
 * yields (i) tests/golden/reference_colliders.json -- every primitive pair function of collision_primitive_core.py --
 * and (ii) tests/golden/pipeline_*.npz -- io.put_model -> make_data -> forward()/step() on nineteen scenes.  The fp64 build
 * of this file reproduces them to 1e-9 (tests/test_oracle_golden_colliders.py, tests/test_oracle_golden_pipeline.py).
 * (iii) The reference's own GJK / EPA known-answer tests (collision_gjk_test.py:307-1002, non-mesh cases) are transcribed in
 * tests/test_oracle_gjk_vectors.py and met by both builds through orc_ccd().
 * Also kept: the reference's hard-coded vectors that need no MuJoCo and physical invariants (tests/test_oracle_*.py).
 * Outside the pin: the MJCF compiler (MuJoCo's C compiler is absent; the fixtures pin step(m, d) given the model arrays).
 *
 * real = double by default (MuJoCo C precision); -DORC_FLOAT builds an fp32 twin used to study
 * fp32 rounding (the reference computes in fp32).
 */
#ifndef MJB_ORACLE_H
#define MJB_ORACLE_H

#ifdef ORC_FLOAT
typedef float real;
#else
typedef double real;
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcModel OrcModel;
typedef struct OrcData OrcData;

OrcModel* orc_model_create(void);
void orc_model_free(OrcModel*);
/* returns 0 on success, -1 if `name` is unknown */
int orc_model_set_int(OrcModel*, const char* name, int value);
int orc_model_set_real(OrcModel*, const char* name, double value);
int orc_model_set_iarr(OrcModel*, const char* name, const int* ptr);
int orc_model_set_rarr(OrcModel*, const char* name, const real* ptr);

OrcData* orc_data_create(int nworld, int nconmax, int njmax);
void orc_data_free(OrcData*);
int orc_data_set_iarr(OrcData*, const char* name, int* ptr);
int orc_data_set_rarr(OrcData*, const char* name, real* ptr);

/* forward dynamics for all worlds (no integration). nthreads<=0: omp default. returns 0 or -1 (unset field) */
int orc_forward(const OrcModel*, OrcData*, int nthreads);
/* forward + integrator (Euler / implicitfast) */
int orc_step(const OrcModel*, OrcData*, int nthreads);
const char* orc_last_error(void);
int orc_sizeof_real(void);

/* exposed math helpers for golden-vector tests (reference math_test.py) */
void orc_closest_segment_point(const real a[3], const real b[3], const real pt[3], real out[3]);
void orc_closest_segment_to_segment_points(const real a0[3], const real a1[3], const real b0[3], const real b1[3], real outa[3], real outb[3]);
int orc_upper_tri_index(int n, int i, int j);
int orc_upper_trid_index(int n, int i, int j);
double orc_halton(int index, int base);
/* solver.py:521-551 shifted cost / gradient / curvature of one elliptic contact at step alpha (quad = (q0, q1, q2), quad1 = (u0, v0, uu),
 * quad2 = (uv, vv, dm)); orc_elliptic_zero: the absolute value at alpha = 0 (solver.py:308-320) */
void orc_elliptic_eval_pt(double alpha, const real* quad, const real* quad1, const real* quad2, double mu, real* out);
void orc_elliptic_zero(const real* quad, const real* quad1, const real* quad2, double mu, real* out);
/* convex pair (GJK / EPA / box multi-contact) on two posed geoms; returns the contact count, witnesses in w1 / w2 (4 x 3 each) */
int orc_ccd(int type1, const real* size1, const real* pos1, const real* mat1, int type2, const real* size2, const real* pos2, const real* mat2,
            real margin, real tolerance, real cutoff, int iterations, int multi, real* dist, real* w1, real* w2, int* overflow);

#ifdef __cplusplus
}
#endif
#endif
