// Test-only: compiles the solver's scalar line-search header (mujoco_warp_b200/csrc/mjb_linesearch.cuh) as plain host C++, so that the
// device source of the per-row / per-contact cost evaluation runs on the CPU against the reference's known-answer vectors
// (solver_test.py:296-350) and against the oracle.  Nothing in the product path uses this file.
#include <cuda_runtime.h>
#include <math.h>
#include <algorithm>
#ifndef __noinline__
#define __noinline__
#endif
using std::max;
using std::min;
static inline float __shfl_xor_sync(unsigned, float v, int) { return v; }
static inline int __shfl_xor_sync(unsigned, int v, int) { return v; }
static inline int __shfl_up_sync(unsigned, int v, int) { return v; }
static inline float __shfl_sync(unsigned, float v, int) { return v; }
static inline int __shfl_sync(unsigned, int v, int) { return v; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
#include "../../mujoco_warp_b200/csrc/mjb_linesearch.cuh"

// solver.py:521-551 _compute_efc_eval_pt_elliptic for the primary row of an elliptic contact: (cost(alpha) - cost(0), grad, hess)
extern "C" void hls_elliptic_eval_pt(float alpha, const float* quad, const float* quad1, const float* quad2, float mu, float* out) {
  EllQ q;
  q.q0 = quad[0]; q.q1 = quad[1]; q.q2 = quad[2]; q.u0 = quad1[0]; q.v0 = quad1[1]; q.uu = quad1[2]; q.uv = quad2[0]; q.vv = quad2[1]; q.dm = quad2[2];
  const P3 p = ell_shifted(mu, q, ell_reference(mu, q), alpha);
  out[0] = p.c; out[1] = p.g; out[2] = p.h;
}
// absolute value at alpha = 0 (solver.py:308-320)
extern "C" void hls_elliptic_zero(const float* quad, const float* quad1, const float* quad2, float mu, float* out) {
  EllQ q;
  q.q0 = quad[0]; q.q1 = quad[1]; q.q2 = quad[2]; q.u0 = quad1[0]; q.v0 = quad1[1]; q.uu = quad1[2]; q.uv = quad2[0]; q.vv = quad2[1]; q.dm = quad2[2];
  const P3 p = ell_zero(mu, q);
  out[0] = p.c; out[1] = p.g; out[2] = p.h;
}
// one row of kind equality (r < ne), friction loss (ne <= r < ne + nf) or inequality, shifted / absolute at alpha = 0
extern "C" void hls_eval_row(int r, float alpha, int ne, int nf, float D, float f, float jaref, float jv, float* out) {
  const P3 p = eval_row(r, alpha, ne, nf, D, f, jaref, jv);
  out[0] = p.c; out[1] = p.g; out[2] = p.h;
}
extern "C" void hls_eval_row_zero(int r, int ne, int nf, float D, float f, float jaref, float jv, float* out) {
  const P3 p = eval_row_zero(r, ne, nf, D, f, jaref, jv);
  out[0] = p.c; out[1] = p.g; out[2] = p.h;
}
