// Test-only: compiles the CUDA convex-collision header (mujoco_warp_b200/csrc/mjb_ccd.cuh) as plain host C++ so that the per-pair device
// routines (GJK, EPA, multi-contact -- scalar code, one lane per geom pair) can be exercised without a GPU and compared with the oracle.
// Nothing in the product path uses this file; the GPU tests run the same routines inside k_collision.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#ifndef __noinline__
#define __noinline__
#endif
using std::max;
using std::min;
// warp intrinsics referenced by helpers in mjb_math.cuh that the pair routines never call
static inline float __shfl_xor_sync(unsigned, float v, int) { return v; }
static inline int __shfl_xor_sync(unsigned, int v, int) { return v; }
static inline int __shfl_up_sync(unsigned, int v, int) { return v; }
static inline float __shfl_sync(unsigned, float v, int) { return v; }
static inline int __shfl_sync(unsigned, int v, int) { return v; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
#include "../../mujoco_warp_b200/csrc/mjb_ccd.cuh"

// same shape as the oracle's orc_ccd (oracle/oracle.c), fp32
extern "C" int hccd_pair(int type1, const float* size1, const float* pos1, const float* mat1, int type2, const float* size2, const float* pos2,
                         const float* mat2, float margin, float tolerance, float cutoff, int gjk_iterations, int epa_iterations, float* dist, float* w1,
                         float* w2, int* overflow) {
  CGeom a, b;
  memset(&a, 0, sizeof a); memset(&b, 0, sizeof b);
  a.pos = ld3(pos1); a.rot = mat1; a.size = ld3(size1); a.margin = margin; a.type = type1;
  b.pos = ld3(pos2); b.rot = mat2; b.size = ld3(size2); b.margin = margin; b.type = type2;
#if CCD_MESH
  a.index = b.index = -1;
#endif
  const int it = std::max(gjk_iterations, epa_iterations);
  float* scratch = new float[ccd_scratch_words(it) + 64]();
  v3 x1[4], x2[4];
  memset(x1, 0, sizeof x1); memset(x2, 0, sizeof x2);
  bool ovf = false;
  const int n = ccd_pair(tolerance, cutoff, gjk_iterations, epa_iterations, a, b, scratch, dist, x1, x2, &ovf);
  for (int k = 0; k < 4; k++) { st3(w1 + 3 * k, x1[k]); st3(w2 + 3 * k, x2[k]); }
  *overflow = ovf ? 1 : 0;
  delete[] scratch;
  return n;
}

#if CCD_MESH
// descriptor entry, same layout as the oracle's OrcGeomDesc / orc_ccd_desc (fp32): mesh tables already offset to the geom's mesh
struct HGeomDesc {
  int type, vertnum, polynum, pad;
  const float *size, *pos, *mat, *vert, *polynormal;
  const int *graph, *polyvertadr, *polyvertnum, *polyvert, *polymapadr, *polymapnum, *polymap;
};
static CGeom from_desc(const HGeomDesc* d, float margin) {
  CGeom c;
  memset(&c, 0, sizeof c);
  c.pos = ld3(d->pos); c.rot = d->mat; c.size = ld3(d->size); c.margin = margin; c.type = d->type;
  c.index = -1; c.vertnum = d->vertnum; c.polynum = d->polynum; c.vert = d->vert; c.polynormal = d->polynormal; c.graph = d->graph;
  c.polyvertadr = d->polyvertadr; c.polyvertnum = d->polyvertnum; c.polyvert = d->polyvert;
  c.polymapadr = d->polymapadr; c.polymapnum = d->polymapnum; c.polymap = d->polymap;
  return c;
}
extern "C" int hccd_desc(const HGeomDesc* d1, const HGeomDesc* d2, float margin, float tolerance, float cutoff, int gjk_iterations, int epa_iterations,
                         float* dist, float* w1, float* w2, int* overflow) {
  const CGeom a = from_desc(d1, margin), b = from_desc(d2, margin);
  const int it = std::max(gjk_iterations, epa_iterations);
  float* scratch = new float[ccd_scratch_words(it) + 64]();
  v3 x1[4], x2[4];
  memset(x1, 0, sizeof x1); memset(x2, 0, sizeof x2);
  bool ovf = false;
  const int n = ccd_pair(tolerance, cutoff, gjk_iterations, epa_iterations, a, b, scratch, dist, x1, x2, &ovf);
  for (int k = 0; k < 4; k++) { st3(w1 + 3 * k, x1[k]); st3(w2 + 3 * k, x2[k]); }
  *overflow = ovf ? 1 : 0;
  delete[] scratch;
  return n;
}
extern "C" void hplane_mesh(const float* n_world, const float* plane_pos, const HGeomDesc* d, float* dist, float* pos) {
  const CGeom c = from_desc(d, 0.f);
  v3 p4[4];
  plane_mesh(ld3(n_world), ld3(plane_pos), c, dist, p4);
  for (int k = 0; k < 4; k++) st3(pos + 3 * k, p4[k]);
}
#endif
