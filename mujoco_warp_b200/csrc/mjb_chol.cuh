// mjb_chol.cuh -- warp-cooperative dense Cholesky in shared memory.
// Replaces the reference's wp.tile_cholesky_inplace / tile_cholesky_solve (cuSolverDx via libmathdx) call sites:
// smooth.py:1308,3151,3256-3260 (inertia blocks) and solver.py:2595-2597 (Newton Hessian).
// A is n x n with leading dimension ld (choose ld odd so that lane-strided row access is bank-conflict free);
// only the lower triangle is read/written.  One warp, lanes map to rows, left-looking column sweep.
#pragma once
#include "mjb_math.cuh"

// In-place lower Cholesky: A = L L^T.  Returns (via *min_diag) the smallest pivot before sqrt for diagnostics.
__device__ __forceinline__ void warp_cholesky(float* A, int n, int ld, int lane) {
  for (int j = 0; j < n; j++) {
    const float* rj = A + j * ld;
    float s = 0.f;
    for (int k = lane; k < j; k += 32) s += rj[k] * rj[k];
    s = warp_sum(s);
    const float ljj = sqrtf(fmaxf(rj[j] - s, MJ_MINVAL));
    const float inv = 1.0f / ljj;
    for (int i = j + 1 + lane; i < n; i += 32) {
      const float* ri = A + i * ld;
      float t = ri[j];
      for (int k = 0; k < j; k++) t -= ri[k] * rj[k];
      A[i * ld + j] = t * inv;
    }
    __syncwarp();
    if (lane == 0) A[j * ld + j] = ljj;
    __syncwarp();
  }
}

// x <- (L L^T)^-1 x, x in shared memory (length n).
__device__ __forceinline__ void warp_chol_solve(const float* L, int n, int ld, float* x, int lane) {
  for (int j = 0; j < n; j++) {  // forward: L y = b
    const float yj = x[j] / L[j * ld + j];
    __syncwarp();
    for (int i = j + 1 + lane; i < n; i += 32) x[i] -= L[i * ld + j] * yj;
    if (lane == 0) x[j] = yj;
    __syncwarp();
  }
  for (int j = n - 1; j >= 0; j--) {  // backward: L^T x = y
    const float xj = x[j] / L[j * ld + j];
    __syncwarp();
    for (int i = lane; i < j; i += 32) x[i] -= L[j * ld + i] * xj;
    if (lane == 0) x[j] = xj;
    __syncwarp();
  }
}
