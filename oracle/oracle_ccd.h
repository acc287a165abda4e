/* oracle_ccd.h -- TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's general convex collision path for analytic
 * convex geoms (sphere, capsule, ellipsoid, cylinder, box; no mesh / hfield, no multi-contact):
 * /root/reference/mujoco_warp/_src/collision_gjk.py :115 support, :281-594 distance sub-algorithm (S1D / S2D / S3D),
 * :635 gjk, :1021-1286 polytope construction, :1319 _epa, :947 _epa_witness, :2303 _inflate, :2350 gjk_phase, :2421 epa_phase;
 * driver collision_convex.py:739-968 (eval_ccd_write_contact).  Included by oracle.c (single translation unit). */

#define CCD_FLOAT_MAX ((real)1e30)
#define CCD_MINVAL ((real)1e-15)
#define CCD_MIN_DIST2 ((real)1e-10)
#define CCD_MIN_DIST3 ((real)1e-10)
#define CCD_MIN_DIST4 ((real)1e-17)
#define CCD_MIN_EPATOL ((real)1e-7)
#define CCD_MINVAL2 ((real)1e-30)
#define CCD_FACE_TOL ((real)0.99999872) /* cos(0.0016) */
#define CCD_EDGE_TOL ((real)0.00159999931) /* sin(0.0016) */
#define CCD_INTERSECT_TOL ((real)0.0000003)
#define CCD_MAX_EPAFACES 5
#define CCD_MAX_EPAHORIZON 24

/* mesh geoms carry their vertex block, the hull graph (NULL: exhaustive search) and the hull polygon tables (collision_core.py Geom);
 * index = cached support vertex (vertex id for the exhaustive search, hull-local id for the hill climb; -1: none) */
typedef struct {
  real pos[3], rot[9], size[3], margin; int type;
  int index, vertnum, polynum;
  const real *vert, *polynormal;          /* vert / polynormal already offset to this mesh (vertadr / polyadr) */
  const int *graph;                       /* this mesh's graph block */
  const int *polyvertadr, *polyvertnum;   /* offset to this mesh's first polygon; the addresses they hold are global */
  const int *polyvert;                    /* global */
  const int *polymapadr, *polymapnum;     /* offset to this mesh's first vertex; addresses are global */
  const int *polymap;                     /* global */
} CGeom;
#define CCD_FLOAT_MIN ((real)-1e30)
#define CCD_MAXDEG 32    /* hull polygons meeting at one vertex (model nmeshdegmax) */
#define CCD_MAXPOLY 64   /* vertices of one hull polygon (model npolygonmax) */
typedef struct {
  int separated, dim; real dist, x1[3], x2[3];
  real simplex[4][3], simplex1[4][3], simplex2[4][3]; int index1[4], index2[4];
  int cindex1, cindex2; /* cached support vertices of the two geoms when gjk returned */
} GjkResult;
typedef struct {
  int status, nvert, nface, nhorizon, maxvert, maxface;
  real (*vert)[3]; int* vert_index; real center[3];
  unsigned* face; real (*face_pr)[3]; real* face_norm2; int horizon[CCD_MAX_EPAHORIZON];
} Polytope;

static inline real csign(real x) { return x < 0 ? (real)-1 : (real)1; } /* warp: sign(0) = +1 */
static inline void v3sub(const real* a, const real* b, real* o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static inline void v3cpy(real* d, const real* s) { d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; }

/* collision_gjk.py:115 support point of a geom inflated by half its margin; *vidx = box corner / mesh vertex id or -1,
 * *cidx = the index to cache for the next query of this geom */
static void ccd_support(const CGeom* g, const real* dir, real* out, int* vidx, int* cidx) {
  const int cached = g->index; /* read first: callers may pass &g->index as cidx */
  *vidx = -1; *cidx = -1;
  if (g->type == GEOM_SPHERE) { for (int i = 0; i < 3; i++) out[i] = g->pos[i] + (g->size[0] + (real)0.5 * g->margin) * dir[i]; return; }
  real ld[3], res[3] = {0, 0, 0};
  matT_vec3(g->rot, dir, ld);
  if (g->type == GEOM_BOX) {
    real t[3] = {csign(ld[0]), csign(ld[1]), csign(ld[2])};
    for (int i = 0; i < 3; i++) res[i] = t[i] * g->size[i];
    *vidx = (t[0] > 0 ? 1 : 0) + (t[1] > 0 ? 2 : 0) + (t[2] > 0 ? 4 : 0);
  } else if (g->type == GEOM_CAPSULE) {
    for (int i = 0; i < 3; i++) res[i] = ld[i] * g->size[0];
    res[2] += csign(ld[2]) * g->size[1];
  } else if (g->type == GEOM_ELLIPSOID) {
    for (int i = 0; i < 3; i++) res[i] = ld[i] * g->size[i];
    normalize3(res);
    for (int i = 0; i < 3; i++) res[i] *= g->size[i];
  } else if (g->type == GEOM_CYLINDER) {
    real d = (real)sqrt((double)(ld[0] * ld[0] + ld[1] * ld[1]));
    if (d > CCD_MINVAL) { real scl = g->size[0] / d; res[0] = ld[0] * scl; res[1] = ld[1] * scl; }
    res[2] = csign(ld[2]) * g->size[1];
  } else if (g->type == GEOM_MESH) { /* :154-194 */
    real max_dist = CCD_FLOAT_MIN;
    if (!g->graph || g->vertnum < 10) {
      if (cached > -1) { *cidx = cached; max_dist = dot3(g->vert + 3 * cached, ld); v3cpy(res, g->vert + 3 * cached); }
      for (int i = 0; i < g->vertnum; i++) {
        real dd = dot3(g->vert + 3 * i, ld);
        if (dd > max_dist) { max_dist = dd; v3cpy(res, g->vert + 3 * i); *cidx = i; }
      }
      *vidx = *cidx;
    } else {
      const int numvert = g->graph[0], *vert_edgeadr = g->graph + 2, *vert_globalid = g->graph + 2 + numvert, *edge_localid = g->graph + 2 + 2 * numvert;
      int prev = -1, imax = cached > -1 ? cached : 0;
      max_dist = dot3(ld, g->vert + 3 * vert_globalid[imax]);
      while (imax != prev) {
        prev = imax;
        int i = vert_edgeadr[imax], subidx = edge_localid[i];
        while (subidx >= 0) {
          real dd = dot3(ld, g->vert + 3 * vert_globalid[subidx]);
          if (dd > max_dist) { imax = subidx; max_dist = dd; }
          i++; subidx = edge_localid[i];
        }
      }
      *cidx = imax; *vidx = vert_globalid[imax];
      v3cpy(res, g->vert + 3 * *vidx);
    }
  }
  matvec3(g->rot, res, out);
  for (int i = 0; i < 3; i++) out[i] += g->pos[i];
  if (g->margin > 0) for (int i = 0; i < 3; i++) out[i] += dir[i] * ((real)0.5 * g->margin);
}

static inline real det3(const real* a, const real* b, const real* c) { real t[3]; cross3(b, c, t); return dot3(a, t); }
static inline int same_sign(real a, real b) { if (a > 0 && b > 0) return 1; if (a < 0 && b < 0) return -1; return 0; }
static void project_origin_line(const real* v1, const real* v2, real* o) {
  real diff[3]; v3sub(v2, v1, diff);
  real scl = -(dot3(v2, diff) / dot3(diff, diff));
  for (int i = 0; i < 3; i++) o[i] = v2[i] + scl * diff[i];
}
static int project_origin_plane(const real* v1, const real* v2, const real* v3, real* o) {
  real d21[3], d31[3], d32[3], n[3], nv, nn;
  v3sub(v2, v1, d21); v3sub(v3, v1, d31); v3sub(v3, v2, d32);
  o[0] = o[1] = o[2] = 0;
  cross3(d32, d21, n); nv = dot3(n, v2); nn = dot3(n, n);
  if (nn == 0) return 1;
  if (nv != 0 && nn > CCD_MINVAL) { for (int i = 0; i < 3; i++) o[i] = (nv / nn) * n[i]; return 0; }
  cross3(d21, d31, n); nv = dot3(n, v1); nn = dot3(n, n);
  if (nn == 0) return 1;
  if (nv != 0 && nn > CCD_MINVAL) { for (int i = 0; i < 3; i++) o[i] = (nv / nn) * n[i]; return 0; }
  cross3(d31, d32, n); nv = dot3(n, v3); nn = dot3(n, n);
  for (int i = 0; i < 3; i++) o[i] = (nv / nn) * n[i];
  return 0;
}
static void S1D(const real* s1, const real* s2, real* l) {
  real po[3]; project_origin_line(s1, s2, po);
  real mu_max = s1[0] - s2[0]; int index = 0;
  real mu = s1[1] - s2[1]; if (rabs(mu) >= rabs(mu_max)) { mu_max = mu; index = 1; }
  mu = s1[2] - s2[2]; if (rabs(mu) >= rabs(mu_max)) { mu_max = mu; index = 2; }
  real C1 = po[index] - s2[index], C2 = s1[index] - po[index];
  if (same_sign(mu_max, C1) && same_sign(mu_max, C2)) { l[0] = C1 / mu_max; l[1] = C2 / mu_max; return; }
  l[0] = 0; l[1] = 1;
}
static void tri_minors(const real* s1, const real* s2, const real* s3, real* M14, real* M24, real* M34) {
  *M14 = s2[1] * s3[2] - s2[2] * s3[1] - s1[1] * s3[2] + s1[2] * s3[1] + s1[1] * s2[2] - s1[2] * s2[1];
  *M24 = s2[0] * s3[2] - s2[2] * s3[0] - s1[0] * s3[2] + s1[2] * s3[0] + s1[0] * s2[2] - s1[2] * s2[0];
  *M34 = s2[0] * s3[1] - s2[1] * s3[0] - s1[0] * s3[1] + s1[1] * s3[0] + s1[0] * s2[1] - s1[1] * s2[0];
}
/* signed areas of (p, b, c), (p, a, c), (p, a, b) in the 2-D projection that drops the axis with the largest minor */
static real tri_cofactors(const real* s1, const real* s2, const real* s3, const real* p, real C[3]) {
  real M14, M24, M34, Mmax; int x, y;
  tri_minors(s1, s2, s3, &M14, &M24, &M34);
  real mu1 = rabs(M14), mu2 = rabs(M24), mu3 = rabs(M34);
  if (mu1 >= mu2 && mu1 >= mu3) { Mmax = M14; x = 1; y = 2; } else if (mu2 >= mu3) { Mmax = M24; x = 0; y = 2; } else { Mmax = M34; x = 0; y = 1; }
  C[0] = p[x] * s2[y] + p[y] * s3[x] + s2[x] * s3[y] - p[x] * s3[y] - p[y] * s2[x] - s3[x] * s2[y];
  C[1] = p[x] * s3[y] + p[y] * s1[x] + s3[x] * s1[y] - p[x] * s1[y] - p[y] * s3[x] - s1[x] * s3[y];
  C[2] = p[x] * s1[y] + p[y] * s2[x] + s1[x] * s2[y] - p[x] * s2[y] - p[y] * s1[x] - s2[x] * s1[y];
  return Mmax;
}
static void S2D(const real* s1, const real* s2, const real* s3, real* l) {
  real po[3];
  if (project_origin_plane(s1, s2, s3, po)) { real v[2]; S1D(s1, s2, v); l[0] = v[0]; l[1] = v[1]; l[2] = 0; return; }
  real C[3], Mmax = tri_cofactors(s1, s2, s3, po, C);
  int c1 = same_sign(Mmax, C[0]), c2 = same_sign(Mmax, C[1]), c3 = same_sign(Mmax, C[2]);
  if (c1 && c2 && c3) { for (int i = 0; i < 3; i++) l[i] = C[i] / Mmax; return; }
  real dmin = CCD_FLOAT_MAX, sub[2], x[3], d;
  l[0] = l[1] = l[2] = 0;
  if (!c1) { S1D(s2, s3, sub); for (int i = 0; i < 3; i++) x[i] = sub[0] * s2[i] + sub[1] * s3[i]; d = dot3(x, x); l[0] = 0; l[1] = sub[0]; l[2] = sub[1]; dmin = d; }
  if (!c2) { S1D(s1, s3, sub); for (int i = 0; i < 3; i++) x[i] = sub[0] * s1[i] + sub[1] * s3[i]; d = dot3(x, x); if (d < dmin) { l[0] = sub[0]; l[1] = 0; l[2] = sub[1]; dmin = d; } }
  if (!c3) { S1D(s1, s2, sub); for (int i = 0; i < 3; i++) x[i] = sub[0] * s1[i] + sub[1] * s2[i]; d = dot3(x, x); if (d < dmin) { l[0] = sub[0]; l[1] = sub[1]; l[2] = 0; } }
}
static void S3D(const real* s1, const real* s2, const real* s3, const real* s4, real* l) {
  real C41 = -det3(s2, s3, s4), C42 = det3(s1, s3, s4), C43 = -det3(s1, s2, s4), C44 = det3(s1, s2, s3);
  real m_det = C41 + C42 + C43 + C44;
  int c1 = same_sign(m_det, C41), c2 = same_sign(m_det, C42), c3 = same_sign(m_det, C43), c4 = same_sign(m_det, C44);
  if (c1 && c2 && c3 && c4) { l[0] = C41 / m_det; l[1] = C42 / m_det; l[2] = C43 / m_det; l[3] = C44 / m_det; return; }
  real dmin = CCD_FLOAT_MAX, sub[3], x[3], d;
  l[0] = l[1] = l[2] = l[3] = 0;
  if (!c1) { S2D(s2, s3, s4, sub); for (int i = 0; i < 3; i++) x[i] = sub[0] * s2[i] + sub[1] * s3[i] + sub[2] * s4[i]; d = dot3(x, x); l[0] = 0; l[1] = sub[0]; l[2] = sub[1]; l[3] = sub[2]; dmin = d; }
  if (!c2) { S2D(s1, s3, s4, sub); for (int i = 0; i < 3; i++) x[i] = sub[0] * s1[i] + sub[1] * s3[i] + sub[2] * s4[i]; d = dot3(x, x); if (d < dmin) { l[0] = sub[0]; l[1] = 0; l[2] = sub[1]; l[3] = sub[2]; dmin = d; } }
  if (!c3) { S2D(s1, s2, s4, sub); for (int i = 0; i < 3; i++) x[i] = sub[0] * s1[i] + sub[1] * s2[i] + sub[2] * s4[i]; d = dot3(x, x); if (d < dmin) { l[0] = sub[0]; l[1] = sub[1]; l[2] = 0; l[3] = sub[2]; dmin = d; } }
  if (!c4) { S2D(s1, s2, s3, sub); for (int i = 0; i < 3; i++) x[i] = sub[0] * s1[i] + sub[1] * s2[i] + sub[2] * s3[i]; d = dot3(x, x); if (d < dmin) { l[0] = sub[0]; l[1] = sub[1]; l[2] = sub[2]; l[3] = 0; } }
}
static void subdistance(int n, real s[4][3], real* l) {
  l[0] = 1; l[1] = l[2] = l[3] = 0;
  if (n == 4) S3D(s[0], s[1], s[2], s[3], l);
  else if (n == 3) { real t[3]; S2D(s[0], s[1], s[2], t); l[0] = t[0]; l[1] = t[1]; l[2] = t[2]; l[3] = 0; }
  else if (n == 2) { real t[2]; S1D(s[0], s[1], t); l[0] = t[0]; l[1] = t[1]; l[2] = l[3] = 0; }
}
static void linear_combine(int n, const real* l, real m[4][3], real* o) {
  o[0] = o[1] = o[2] = 0;
  for (int k = 0; k < (n < 1 ? 1 : n); k++) for (int i = 0; i < 3; i++) o[i] += l[k] * m[k][i];
}

/* collision_gjk.py:635 (is_discrete = false for analytic geoms: no direction tuning, tolerance-based stop) */
static void ccd_gjk(real tolerance, int iterations, const CGeom* g1_in, const CGeom* g2_in, const real* x1_0, const real* x2_0, real cutoff, int discrete, GjkResult* r) {
  CGeom ga = *g1_in, gb = *g2_in, *g1 = &ga, *g2 = &gb; /* by value: the cached support indices advance with the iterations (:675-679) */
  /* discrete pairs (box / mesh, no margin) converge in finitely many steps: no tolerance (collision_gjk.py:662-663) */
  real lmbda[4] = {1, 0, 0, 0}, x_k[3], epsilon = discrete ? 0 : (real)0.5 * tolerance * tolerance, min_norm = discrete ? CCD_MINVAL : tolerance;
  int n = 0;
  memset(r, 0, sizeof *r);
  v3sub(x1_0, x2_0, x_k);
  real xnorm = (real)sqrt((double)dot3(x_k, x_k)), xnorm_prev = 0;
  for (int it = 0; it < iterations; it++) {
    if (xnorm < min_norm || rabs(xnorm_prev - xnorm) < CCD_MINVAL) break;
    real dir_neg[3] = {x_k[0] / xnorm, x_k[1] / xnorm, x_k[2] / xnorm};
    if (discrete && xnorm < (real)1e-4) { /* :609-627 direction tuning when the search direction is noisy */
      if (n == 2) {
        real edge[3]; v3sub(r->simplex[1], r->simplex[0], edge);
        real en2 = dot3(edge, edge);
        if (en2 > CCD_MINVAL2) {
          real proj = dot3(dir_neg, edge) / en2;
          for (int i = 0; i < 3; i++) dir_neg[i] -= proj * edge[i];
          real dn = len3(dir_neg);
          if (dn > CCD_MINVAL) for (int i = 0; i < 3; i++) dir_neg[i] /= dn;
        }
      } else if (n == 3) {
        real e1[3], e2[3], nrm[3];
        v3sub(r->simplex[1], r->simplex[0], e1); v3sub(r->simplex[2], r->simplex[0], e2); cross3(e1, e2, nrm);
        real nn = len3(nrm);
        if (nn > CCD_MINVAL) { real sg = csign(dot3(dir_neg, nrm)); for (int i = 0; i < 3; i++) dir_neg[i] = sg * nrm[i] / nn; }
      }
    }
    real dpos[3] = {-dir_neg[0], -dir_neg[1], -dir_neg[2]};
    ccd_support(g1, dpos, r->simplex1[n], &r->index1[n], &g1->index);
    ccd_support(g2, dir_neg, r->simplex2[n], &r->index2[n], &g2->index);
    v3sub(r->simplex1[n], r->simplex2[n], r->simplex[n]);
    real dk[3]; v3sub(x_k, r->simplex[n], dk);
    if (dot3(x_k, dk) < epsilon) break;
    real lower = dot3(x_k, r->simplex[n]);
    if (cutoff == 0) { if (lower > 0) { r->separated = 1; r->dim = 0; r->dist = CCD_FLOAT_MAX; r->cindex1 = g1->index; r->cindex2 = g2->index; return; } }
    else if (cutoff < CCD_FLOAT_MAX) { if (lower > 0 && lower >= cutoff * xnorm) { r->separated = 1; r->dim = 0; r->dist = CCD_FLOAT_MAX; r->cindex1 = g1->index; r->cindex2 = g2->index; return; } }
    subdistance(n + 1, r->simplex, lmbda);
    n = 0;
    for (int i = 0; i < 4; i++) {
      if (lmbda[i] == 0) continue;
      v3cpy(r->simplex[n], r->simplex[i]); v3cpy(r->simplex1[n], r->simplex1[i]); v3cpy(r->simplex2[n], r->simplex2[i]);
      r->index1[n] = r->index1[i]; r->index2[n] = r->index2[i]; lmbda[n] = lmbda[i];
      n++;
    }
    if (n < 1) break;
    linear_combine(n, lmbda, r->simplex, x_k);
    xnorm_prev = xnorm;
    xnorm = (real)sqrt((double)dot3(x_k, x_k));
    if (n == 4) break;
  }
  r->separated = 0;
  if (n == 0) { v3cpy(r->x1, x1_0); v3cpy(r->x2, x2_0); } else { linear_combine(n, lmbda, r->simplex1, r->x1); linear_combine(n, lmbda, r->simplex2, r->x2); }
  if (xnorm > 0) {
    real dir[3] = {x_k[0] / xnorm, x_k[1] / xnorm, x_k[2] / xnorm}, nd[3] = {-dir[0], -dir[1], -dir[2]}, p1[3], p2[3], dd[3]; int vi, ci;
    ccd_support(g1, nd, p1, &vi, &ci); ccd_support(g2, dir, p2, &vi, &ci);
    v3sub(p1, p2, dd);
    r->separated = dot3(x_k, dd) > 0;
  }
  r->dist = (n == 4 && !r->separated) ? 0 : xnorm;
  r->dim = n;
  r->cindex1 = g1->index; r->cindex2 = g2->index;
}

/* ---- EPA */
static int same_side(const real* p0, const real* p1, const real* p2, const real* p3) {
  real a[3], b[3], n[3], c[3], neg[3] = {-p0[0], -p0[1], -p0[2]};
  v3sub(p1, p0, a); v3sub(p2, p0, b); cross3(a, b, n); v3sub(p3, p0, c);
  real d1 = dot3(n, c), d2 = dot3(n, neg);
  return (d1 > 0 && d2 > 0) || (d1 < 0 && d2 < 0);
}
static int test_tetra(const real* p0, const real* p1, const real* p2, const real* p3) {
  return same_side(p0, p1, p2, p3) && same_side(p1, p2, p3, p0) && same_side(p2, p3, p0, p1) && same_side(p3, p0, p1, p2);
}
static void tri_affine_coord(const real* v1, const real* v2, const real* v3, const real* p, real* l) {
  real C[3], Mmax = tri_cofactors(v1, v2, v3, p, C);
  for (int i = 0; i < 3; i++) l[i] = C[i] / Mmax;
}
static int tri_point_intersect(const real* v1, const real* v2, const real* v3, const real* p) {
  real l[3]; tri_affine_coord(v1, v2, v3, p, l);
  if (l[0] < 0 || l[1] < 0 || l[2] < 0) return 0;
  real pr[3], d[3];
  for (int i = 0; i < 3; i++) pr[i] = v1[i] * l[0] + v2[i] * l[1] + v3[i] * l[2];
  v3sub(pr, p, d);
  return len3(d) < CCD_MINVAL;
}
static void pt_mink(const Polytope* pt, int v, real* o) { v3sub(pt->vert[2 * v], pt->vert[2 * v + 1], o); }
static real attach_face(Polytope* pt, int idx, int v1, int v2, int v3) {
  if (pt->nface == pt->maxface) return 0;
  real p1[3], p2[3], p3[3], r[3], d[3];
  pt_mink(pt, v1, p1); pt_mink(pt, v2, p2); pt_mink(pt, v3, p3);
  if (project_origin_plane(p3, p2, p1, r)) return 0;
  v3sub(p1, pt->center, d);
  if (dot3(r, d) < 0) for (int i = 0; i < 3; i++) r[i] = -r[i];
  pt->face[idx] = (unsigned)(v1 + (v2 << 10) + (v3 << 20));
  v3cpy(pt->face_pr[idx], r);
  pt->face_norm2[idx] = dot3(r, r);
  return pt->face_norm2[idx];
}
static void epa_support_c(Polytope* pt, int idx, const CGeom* g1, const CGeom* g2, const real* dir, int* c1, int* c2) {
  real nd[3] = {-dir[0], -dir[1], -dir[2]};
  ccd_support(g1, dir, pt->vert[2 * idx], &pt->vert_index[2 * idx], c1);
  ccd_support(g2, nd, pt->vert[2 * idx + 1], &pt->vert_index[2 * idx + 1], c2);
}
static void epa_support(Polytope* pt, int idx, const CGeom* g1, const CGeom* g2, const real* dir) { int c1, c2; epa_support_c(pt, idx, g1, g2, dir, &c1, &c2); }
static void replace_simplex3(const Polytope* pt, int v1, int v2, int v3, GjkResult* r) {
  int v[3] = {v1, v2, v3};
  for (int k = 0; k < 3; k++) {
    v3cpy(r->simplex1[k], pt->vert[2 * v[k]]); v3cpy(r->simplex2[k], pt->vert[2 * v[k] + 1]);
    v3sub(r->simplex1[k], r->simplex2[k], r->simplex[k]);
    r->index1[k] = pt->vert_index[2 * v[k]]; r->index2[k] = pt->vert_index[2 * v[k] + 1];
  }
}
static int ray_triangle(const real* v1, const real* v2, const real* v3, const real* v4, const real* v5) {
  real a[3], b[3], c[3], d[3];
  v3sub(v3, v1, a); v3sub(v4, v1, b); v3sub(v5, v1, c); v3sub(v2, v1, d);
  real vol1 = det3(a, b, d), vol2 = det3(b, c, d), vol3 = det3(c, a, d);
  if (vol1 >= 0 && vol2 >= 0 && vol3 >= 0) return 1;
  if (vol1 <= 0 && vol2 <= 0 && vol3 <= 0) return -1;
  return 0;
}
static void load_simplex(Polytope* pt, const GjkResult* r, int n) {
  for (int k = 0; k < n; k++) { v3cpy(pt->vert[2 * k], r->simplex1[k]); v3cpy(pt->vert[2 * k + 1], r->simplex2[k]); pt->vert_index[2 * k] = r->index1[k]; pt->vert_index[2 * k + 1] = r->index2[k]; }
}
/* :1021 hexahedron from a 1-simplex; status -1 = fall back to the 2-simplex written into r */
static void polytope2(Polytope* pt, GjkResult* r, const CGeom* g1, const CGeom* g2) {
  real diff[3]; v3sub(r->simplex[1], r->simplex[0], diff);
  for (int i = 0; i < 3; i++) pt->center[i] = (real)0.5 * (r->simplex[0][i] + r->simplex[1][i]);
  real value = CCD_FLOAT_MAX; int index = 0;
  for (int i = 0; i < 3; i++) if (rabs(diff[i]) < value) { value = rabs(diff[i]); index = i; }
  real e[3] = {0, 0, 0}, d1[3], d2[3], d3[3], R[9];
  e[index] = 1;
  cross3(e, diff, d1);
  { real n = len3(diff), u1 = diff[0] / n, u2 = diff[1] / n, u3 = diff[2] / n, s = (real)0.86602540378, c = (real)-0.5; /* :885 rotation by 120 deg */
    R[0] = c + u1 * u1 * (1 - c); R[1] = u1 * u2 * (1 - c) - u3 * s; R[2] = u1 * u3 * (1 - c) + u2 * s;
    R[3] = u2 * u1 * (1 - c) + u3 * s; R[4] = c + u2 * u2 * (1 - c); R[5] = u2 * u3 * (1 - c) - u1 * s;
    R[6] = u1 * u3 * (1 - c) - u2 * s; R[7] = u2 * u3 * (1 - c) + u1 * s; R[8] = c + u3 * u3 * (1 - c); }
  matvec3(R, d1, d2); matvec3(R, d2, d3);
  load_simplex(pt, r, 2);
  real t[3], nn;
  nn = len3(d1); for (int i = 0; i < 3; i++) t[i] = d1[i] / nn; epa_support(pt, 2, g1, g2, t);
  nn = len3(d2); for (int i = 0; i < 3; i++) t[i] = d2[i] / nn; epa_support(pt, 3, g1, g2, t);
  nn = len3(d3); for (int i = 0; i < 3; i++) t[i] = d3[i] / nn; epa_support(pt, 4, g1, g2, t);
  static const int F[6][3] = {{0, 2, 3}, {0, 4, 2}, {0, 3, 4}, {1, 3, 2}, {1, 2, 4}, {1, 4, 3}};
  for (int f = 0; f < 6; f++)
    if (attach_face(pt, f, F[f][0], F[f][1], F[f][2]) < CCD_MIN_DIST2) { pt->status = -1; replace_simplex3(pt, F[f][0], F[f][1], F[f][2], r); return; }
  real v2[3], v3[3], v4[3];
  pt_mink(pt, 2, v2); pt_mink(pt, 3, v3); pt_mink(pt, 4, v4);
  if (!ray_triangle(r->simplex[0], r->simplex[1], v2, v3, v4)) { pt->status = 1; return; }
  pt->nvert = 5; pt->nface = 6; pt->status = 0;
}
/* :1114 hexahedron from a 2-simplex */
static void polytope3(Polytope* pt, const GjkResult* r, const CGeom* g1, const CGeom* g2) {
  for (int i = 0; i < 3; i++) pt->center[i] = (r->simplex[0][i] + r->simplex[1][i] + r->simplex[2][i]) * (real)(1.0 / 3.0);
  real a[3], b[3], n[3];
  v3sub(r->simplex[1], r->simplex[0], a); v3sub(r->simplex[2], r->simplex[0], b); cross3(a, b, n);
  real norm = len3(n);
  if (norm < CCD_MINVAL) { pt->status = 2; return; }
  for (int i = 0; i < 3; i++) n[i] /= norm;
  load_simplex(pt, r, 3);
  real nn[3] = {-n[0], -n[1], -n[2]};
  epa_support(pt, 3, g1, g2, nn);
  epa_support(pt, 4, g1, g2, n);
  const real *v1 = r->simplex[0], *v2 = r->simplex[1], *v3 = r->simplex[2];
  real v4[3], v5[3];
  pt_mink(pt, 3, v4); pt_mink(pt, 4, v5);
  if (tri_point_intersect(v1, v2, v3, v4)) { pt->status = 3; return; }
  if (tri_point_intersect(v1, v2, v3, v5)) { pt->status = 4; return; }
  if (r->dist > (real)1e-5 && !test_tetra(v1, v2, v3, v4) && !test_tetra(v1, v2, v3, v5)) { pt->status = 5; return; }
  static const int F[6][3] = {{4, 0, 1}, {4, 2, 0}, {4, 1, 2}, {3, 1, 0}, {3, 0, 2}, {3, 2, 1}};
  for (int f = 0; f < 6; f++) if (attach_face(pt, f, F[f][0], F[f][1], F[f][2]) < CCD_MIN_DIST3) { pt->status = 6 + f; return; }
  pt->nvert = 5; pt->nface = 6; pt->status = 0;
}
/* :1207 tetrahedron from a 3-simplex */
static void polytope4(Polytope* pt, GjkResult* r) {
  for (int i = 0; i < 3; i++) pt->center[i] = (real)0.25 * (r->simplex[0][i] + r->simplex[1][i] + r->simplex[2][i] + r->simplex[3][i]);
  load_simplex(pt, r, 4);
  static const int F[4][3] = {{0, 1, 2}, {0, 3, 1}, {0, 2, 3}, {3, 2, 1}};
  real dist[4]; int idx = 0;
  for (int f = 0; f < 4; f++) {
    dist[f] = attach_face(pt, f, F[f][0], F[f][1], F[f][2]);
    if (dist[f] < CCD_MIN_DIST4) { pt->status = -1; replace_simplex3(pt, F[f][0], F[f][1], F[f][2], r); return; }
    if (f == 1) idx = dist[0] < dist[1] ? 0 : 1;
    else if (f > 1) idx = dist[f] < dist[idx] ? f : idx;
  }
  if (!test_tetra(r->simplex[0], r->simplex[1], r->simplex[2], r->simplex[3])) {
    if (dist[idx] > CCD_MINVAL) { pt->status = 12; return; }
    pt->status = -1; replace_simplex3(pt, F[idx][0], F[idx][1], F[idx][2], r); return;
  }
  pt->nvert = 4; pt->nface = 4; pt->status = 0;
}
#define FACE_DELETED 0x80000000u
#define FACE_INVALID 0x40000000u
static int add_edge(Polytope* pt, int e1, int e2) {
  int n = pt->nhorizon;
  if (n < 0) return -1;
  int edge = ((e1 < e2 ? e1 : e2) << 10) | (e1 < e2 ? e2 : e1);
  for (int i = 0; i < n; i++) if (edge == pt->horizon[i]) { pt->horizon[i] = pt->horizon[n - 1]; return n - 1; }
  if (n == CCD_MAX_EPAHORIZON) return -1;
  pt->horizon[n] = edge;
  return n + 1;
}
/* :1319; returns the index of the closest face (or -1) and writes the witness points / distance */
static int ccd_epa(real tolerance, int iterations, Polytope* pt, const CGeom* g1_in, const CGeom* g2_in, int discrete, real* dist, real* x1, real* x2, int* ovf) {
  CGeom ga = *g1_in, gb = *g2_in, *g1 = &ga, *g2 = &gb; /* :1372-1373 the cached indices follow the expansion */
  real upper = CCD_FLOAT_MAX, upper2 = CCD_FLOAT_MAX, epsilon = discrete ? CCD_MIN_EPATOL : tolerance;
  int idx = -1, pidx = -1, nvalid = pt->nface;
  if (iterations > 1000) iterations = 1000;
  for (int it = 0; it < iterations; it++) {
    pidx = idx; idx = -1;
    real lower2 = CCD_FLOAT_MAX;
    for (int i = 0; i < pt->nface; i++) if (!(pt->face[i] & (FACE_DELETED | FACE_INVALID)) && pt->face_norm2[i] < lower2) { idx = i; lower2 = pt->face_norm2[i]; }
    if (lower2 > upper2 || idx < 0) { idx = pidx; break; }
    if (lower2 <= 0) break;
    real lower = (real)sqrt((double)lower2), fp[3], dir[3], w[3];
    int wi = pt->nvert;
    v3cpy(fp, pt->face_pr[idx]);
    for (int i = 0; i < 3; i++) dir[i] = fp[i] / lower;
    epa_support_c(pt, wi, g1, g2, dir, &g1->index, &g2->index);
    pt_mink(pt, wi, w);
    pt->nvert++;
    real upper_k = dot3(fp, w) / lower;
    if (upper_k < upper) { upper = upper_k; upper2 = upper * upper; }
    if (upper - lower < epsilon) break;
    if (discrete) { /* :1377-1385 a repeated support point ends the expansion */
      int rep = 0;
      for (int i = 0; i < pt->nvert - 1; i++) if (pt->vert_index[2 * i] == pt->vert_index[2 * wi] && pt->vert_index[2 * i + 1] == pt->vert_index[2 * wi + 1]) { rep = 1; break; }
      if (rep) break;
    }
    nvalid--;
    pt->face[idx] |= FACE_DELETED;
    { unsigned f = pt->face[idx]; int a = f & 0x3FF, b = (f >> 10) & 0x3FF, c = (f >> 20) & 0x3FF;
      pt->nhorizon = add_edge(pt, a, b); pt->nhorizon = add_edge(pt, b, c); pt->nhorizon = add_edge(pt, c, a); }
    if (pt->nhorizon == -1) { *ovf = 1; idx = -1; break; }
    for (int i = 0; i < pt->nface; i++) {
      if (pt->face[i] & FACE_DELETED) continue;
      if (dot3(pt->face_pr[i], w) - pt->face_norm2[i] > (real)1e-10) {
        if (!(pt->face[i] & (FACE_DELETED | FACE_INVALID))) nvalid--;
        pt->face[i] |= FACE_DELETED;
        unsigned f = pt->face[i]; int a = f & 0x3FF, b = (f >> 10) & 0x3FF, c = (f >> 20) & 0x3FF;
        pt->nhorizon = add_edge(pt, a, b); pt->nhorizon = add_edge(pt, b, c); pt->nhorizon = add_edge(pt, c, a);
        if (pt->nhorizon == -1) { *ovf = 1; idx = -1; break; }
      }
    }
    for (int i = 0; i < pt->nhorizon; i++) {
      int e = pt->horizon[i];
      real dist2 = attach_face(pt, pt->nface, wi, e & 0x3FF, (e >> 10) & 0x3FF);
      if (dist2 == 0) { idx = -1; break; }
      pt->nface++;
      if (dist2 >= lower2 && dist2 <= upper2) nvalid++; else pt->face[pt->nface - 1] |= FACE_INVALID;
    }
    if (nvalid == 0 || idx == -1) break;
    pt->nhorizon = 0;
  }
  if (idx > -1) { /* :947 witness points from the barycentric coordinates of the projection on the closest face */
    unsigned f = pt->face[idx]; int a = f & 0x3FF, b = (f >> 10) & 0x3FF, c = (f >> 20) & 0x3FF;
    real v1[3], v2[3], v3[3], l[3];
    pt_mink(pt, a, v1); pt_mink(pt, b, v2); pt_mink(pt, c, v3);
    tri_affine_coord(v1, v2, v3, pt->face_pr[idx], l);
    for (int i = 0; i < 3; i++) {
      x2[i] = pt->vert[2 * a + 1][i] * l[0] + pt->vert[2 * b + 1][i] * l[1] + pt->vert[2 * c + 1][i] * l[2];
      x1[i] = pt->vert[2 * a][i] * l[0] + pt->vert[2 * b][i] * l[1] + pt->vert[2 * c][i] * l[2];
    }
    *dist = -(real)sqrt((double)pt->face_norm2[idx]);
    return idx;
  }
  *dist = 0; x1[0] = x1[1] = x1[2] = x2[0] = x2[1] = x2[2] = 0;
  return -1;
}

/* ---- multi-contact recovery for box pairs (collision_gjk.py:1503 _feature_dim, :1703-1888 box normals / edges / faces,
 * :1916-2056 polygon clipping, :2076 multicontact; mesh branches omitted) */
static int feature_dim(const Polytope* pt, const int face[3], int offset, int fidx[3], real fvert[3][3]) {
  int v1i = pt->vert_index[2 * face[0] + offset], v2i = pt->vert_index[2 * face[1] + offset], v3i = pt->vert_index[2 * face[2] + offset];
  fidx[0] = v1i; fidx[1] = v2i; fidx[2] = v3i;
  for (int k = 0; k < 3; k++) v3cpy(fvert[k], pt->vert[2 * face[k] + offset]);
  if (v1i != v2i) return (v3i == v1i || v3i == v2i) ? 2 : 3;
  fidx[1] = v3i; v3cpy(fvert[1], pt->vert[2 * face[2] + offset]);
  return v1i != v3i ? 2 : 1;
}
static int box_normals2(const real* mat, const real* n, real nout[3][3], int* iout) {
  static const real FN[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  real ln[3]; matT_vec3(mat, n, ln); normalize3(ln);
  for (int i = 0; i < 6; i++) if (dot3(ln, FN[i]) > CCD_FACE_TOL) { matvec3(mat, FN[i], nout[0]); iout[0] = i; return 1; }
  return 0;
}
static int box_normals(int fdim, const int* fi, const real* mat, const real* dir, real nout[3][3], int* iout) {
  int v1 = fi[0], v2 = fi[1], v3 = fi[2];
  if (fdim == 3) {
    int c = 0;
    real x = (real)(((v1 & 1) && (v2 & 1) && (v3 & 1)) ? 1 : 0) - (real)((!(v1 & 1) && !(v2 & 1) && !(v3 & 1)) ? 1 : 0);
    real y = (real)(((v1 & 2) && (v2 & 2) && (v3 & 2)) ? 1 : 0) - (real)((!(v1 & 2) && !(v2 & 2) && !(v3 & 2)) ? 1 : 0);
    real z = (real)(((v1 & 4) && (v2 & 4) && (v3 & 4)) ? 1 : 0) - (real)((!(v1 & 4) && !(v2 & 4) && !(v3 & 4)) ? 1 : 0);
    real l[3] = {x, y, z}; matvec3(mat, l, nout[0]);
    real sgn = x + y + z;
    if (x != 0) iout[c++] = 0;
    if (y != 0) iout[c++] = 2;
    if (z != 0) iout[c++] = 4;
    if (sgn == -1) iout[0] = iout[0] + 1;
    if (c == 1) return 1;
    return box_normals2(mat, dir, nout, iout);
  }
  if (fdim == 2) {
    int c = 0;
    real x = (real)(((v1 & 1) && (v2 & 1)) ? 1 : 0) - (real)((!(v1 & 1) && !(v2 & 1)) ? 1 : 0);
    real y = (real)(((v1 & 2) && (v2 & 2)) ? 1 : 0) - (real)((!(v1 & 2) && !(v2 & 2)) ? 1 : 0);
    real z = (real)(((v1 & 4) && (v2 & 4)) ? 1 : 0) - (real)((!(v1 & 4) && !(v2 & 4)) ? 1 : 0);
    if (x != 0) { real l[3] = {x, 0, 0}; matvec3(mat, l, nout[c]); iout[c] = x > 0 ? 0 : 1; c++; }
    if (y != 0) { real l[3] = {0, y, 0}; matvec3(mat, l, nout[c]); iout[c] = y > 0 ? 2 : 3; c++; }
    if (z != 0) { real l[3] = {0, 0, z}; matvec3(mat, l, nout[c]); iout[c] = z > 0 ? 4 : 5; c++; }
    if (c == 1 || c == 2) return c;
    return box_normals2(mat, dir, nout, iout);
  }
  if (fdim == 1) {
    real x = (v1 & 1) ? (real)1 : (real)-1, y = (v1 & 2) ? (real)1 : (real)-1, z = (v1 & 4) ? (real)1 : (real)-1;
    real lx[3] = {x, 0, 0}, ly[3] = {0, y, 0}, lz[3] = {0, 0, z};
    matvec3(mat, lx, nout[0]); matvec3(mat, ly, nout[1]); matvec3(mat, lz, nout[2]);
    iout[0] = x > 0 ? 0 : 1; iout[1] = y > 0 ? 2 : 3; iout[2] = z > 0 ? 4 : 5;
    return 3;
  }
  return 0;
}
static int box_edge_normals(int dim, const CGeom* g, const real* v1, const real* v2, int v1i, real nout[3][3], real endvert[3][3]) {
  if (dim == 2) { v3cpy(endvert[0], v2); v3sub(v2, v1, nout[0]); normalize3(nout[0]); return 1; }
  if (dim == 1) {
    real x = (v1i & 1) ? g->size[0] : -g->size[0], y = (v1i & 2) ? g->size[1] : -g->size[1], z = (v1i & 4) ? g->size[2] : -g->size[2];
    real l[3][3] = {{-x, y, z}, {x, -y, z}, {x, y, -z}};
    for (int k = 0; k < 3; k++) {
      matvec3(g->rot, l[k], endvert[k]);
      for (int i = 0; i < 3; i++) endvert[k][i] += g->pos[i];
      v3sub(endvert[k], v1, nout[k]); normalize3(nout[k]);
    }
    return 3;
  }
  return 0;
}
static int box_face(const CGeom* g, int idx, real face[4][3]) {
  const real sx = g->size[0], sy = g->size[1], sz = g->size[2];
  const real L[6][4][3] = {
    {{sx, sy, sz}, {sx, sy, -sz}, {sx, -sy, -sz}, {sx, -sy, sz}}, {{-sx, sy, -sz}, {-sx, sy, sz}, {-sx, -sy, sz}, {-sx, -sy, -sz}},
    {{-sx, sy, -sz}, {sx, sy, -sz}, {sx, sy, sz}, {-sx, sy, sz}}, {{-sx, -sy, sz}, {sx, -sy, sz}, {sx, -sy, -sz}, {-sx, -sy, -sz}},
    {{-sx, sy, sz}, {sx, sy, sz}, {sx, -sy, sz}, {-sx, -sy, sz}}, {{sx, sy, -sz}, {-sx, sy, -sz}, {-sx, -sy, -sz}, {sx, -sy, -sz}}};
  if (idx < 0 || idx > 5) return 0;
  for (int k = 0; k < 4; k++) { matvec3(g->rot, L[idx][k], face[k]); for (int i = 0; i < 3; i++) face[k][i] += g->pos[i]; }
  return 4;
}
static real area4(const real* a, const real* b, const real* c, const real* d) {
  real ad[3], db[3], bc[3], ca[3], c1[3], c2[3], s[3];
  v3sub(a, d, ad); v3sub(d, b, db); v3sub(b, c, bc); v3sub(c, a, ca);
  cross3(ad, db, c1); cross3(bc, ca, c2);
  for (int i = 0; i < 3; i++) s[i] = c1[i] + c2[i];
  return (real)0.5 * len3(s);
}
static void polygon_quad(real poly[][3], int np, int res[4]) { /* :1463 maximum-area quadrilateral of a convex polygon */
  int b = 1, c = 2, d = 3;
  res[0] = 0; res[1] = b; res[2] = c; res[3] = d;
  real m = area4(poly[0], poly[b], poly[c], poly[d]);
  for (int a = 0; a < np; a++) {
    for (;;) {
      real mn = area4(poly[a], poly[b], poly[c], poly[(d + 1) % np]);
      if (mn <= m) break;
      m = mn; d = (d + 1) % np; res[0] = a; res[1] = b; res[2] = c; res[3] = d;
      for (;;) {
        mn = area4(poly[a], poly[b], poly[(c + 1) % np], poly[d]);
        if (mn <= m) break;
        m = mn; c = (c + 1) % np; res[0] = a; res[1] = b; res[2] = c; res[3] = d;
      }
      for (;;) {
        mn = area4(poly[a], poly[(b + 1) % np], poly[c], poly[d]);
        if (mn <= m) break;
        m = mn; b = (b + 1) % np; res[0] = a; res[1] = b; res[2] = c; res[3] = d;
      }
    }
    if (b == a) { b = (b + 1) % np; if (c == b) { c = (c + 1) % np; if (d == c) d = (d + 1) % np; } }
  }
}
/* :1941 clip polygon face2 against the side planes of face1 (extruded along n); witness2 on the clipped polygon, witness1 = witness2 - dir */
static int polygon_clip(real face1[][3], int nface1, real face2[][3], int nface2, const real* n, const real* dir, real w1[4][3], real w2[4][3]) {
  if (nface1 < 3) return 0;
  real pn[CCD_MAXPOLY][3], pd[CCD_MAXPOLY], bufA[2 * CCD_MAXPOLY][3], bufB[2 * CCD_MAXPOLY][3];
  const int cap = 2 * CCD_MAXPOLY; /* the reference sizes these buffers 2 x the model's largest polygon (collision_convex.py:1229-1236) */
  if (nface1 > CCD_MAXPOLY || nface2 > cap) return 0;
  for (int i = 0; i < nface1; i++) { /* :1916 _plane_normal */
    const real *a = face1[i], *b = face1[(i + 1) % nface1];
    real ba[3], res[3]; v3sub(b, a, ba); cross3(ba, n, res);
    pd[i] = dot3(res, a); v3cpy(pn[i], res);
  }
  real (*poly)[3] = bufA, (*clip)[3] = bufB;
  int np = nface2, nc = 0;
  for (int i = 0; i < nface2; i++) v3cpy(poly[i], face2[i]);
  for (int e = 0; e < nface1; e++) {
    for (int i = 0; i < np; i++) {
      const real *P = poly[i], *Q = poly[(i + 1) % np];
      real dP[3], dQ[3]; v3sub(P, face1[e], dP); v3sub(Q, face1[e], dQ);
      int in1 = dot3(dP, pn[e]) > (real)-1e-10, in2 = dot3(dQ, pn[e]) > (real)-1e-10;
      if (!in1 && !in2) continue;
      if (in1 && in2) { if (nc < cap) v3cpy(clip[nc], Q); nc++; continue; }
      real pq[3]; v3sub(Q, P, pq);
      real dt = dot3(pn[e], pq), t = rabs(dt) < (real)1e-10 ? CCD_FLOAT_MAX : (pd[e] - dot3(pn[e], P)) / dt;
      if (t > -CCD_INTERSECT_TOL && t < 1 + CCD_INTERSECT_TOL) {
        t = rclamp(t, 0, 1);
        if (nc < cap) for (int k = 0; k < 3; k++) clip[nc][k] = P[k] + t * pq[k];
        nc++;
      }
      if (in2) { if (nc < cap) v3cpy(clip[nc], Q); nc++; }
    }
    if (nc > cap) nc = cap;
    real (*tmp)[3] = poly; poly = clip; clip = tmp;
    np = nc; nc = 0;
  }
  if (np < 1) return 0;
  if (nface2 == 2 && np > 2) { /* an edge: keep the two farthest points */
    int b1 = 0, b2 = 1; real maxd = 0;
    for (int i = 0; i < np; i++) for (int j = i + 1; j < np; j++) { real df[3]; v3sub(poly[j], poly[i], df); real d2 = dot3(df, df); if (d2 > maxd) { maxd = d2; b1 = i; b2 = j; } }
    v3cpy(w2[0], poly[b1]); v3sub(w2[0], dir, w1[0]); v3cpy(w2[1], poly[b2]); v3sub(w2[1], dir, w1[1]);
    return 2;
  }
  if (np > 4) {
    int q[4]; polygon_quad(poly, np, q);
    for (int i = 0; i < 4; i++) { v3cpy(w2[i], poly[q[i]]); v3sub(w2[i], dir, w1[i]); }
    return 4;
  }
  for (int i = 0; i < np; i++) { v3cpy(w2[i], poly[i]); v3sub(w2[i], dir, w1[i]); }
  return np;
}
/* collision_gjk.py:1556-1581 common polygon ids of two vertices' polygon lists (at most two) */
static int mesh_intersect(const int* a1, int n1, const int* a2, int n2, int res[2]) {
  int count = 0;
  for (int i = 0; i < n1; i++) for (int j = 0; j < n2; j++) if (a1[i] == a2[j]) { res[count++] = a1[i]; if (count == 2) return 2; }
  return count;
}
/* :1585-1651 possible hull-polygon normals of a mesh feature given by up to three vertices */
static int mesh_normals(int fdim, const int* fi, const CGeom* g, real nout[][3], int* iout) {
  const int *m1 = g->polymap + g->polymapadr[fi[0]], n1 = g->polymapnum[fi[0]];
  if (fdim == 3) {
    int e[2], f[2];
    int n = mesh_intersect(m1, n1, g->polymap + g->polymapadr[fi[1]], g->polymapnum[fi[1]], e);
    if (n == 0) return 0;
    n = mesh_intersect(e, n, g->polymap + g->polymapadr[fi[2]], g->polymapnum[fi[2]], f);
    if (n == 0) return 0;
    matvec3(g->rot, g->polynormal + 3 * f[0], nout[0]); iout[0] = f[0];
    return 1;
  }
  if (fdim == 2) {
    int e[2];
    int n = mesh_intersect(m1, n1, g->polymap + g->polymapadr[fi[1]], g->polymapnum[fi[1]], e);
    for (int i = 0; i < n; i++) { matvec3(g->rot, g->polynormal + 3 * e[i], nout[i]); iout[i] = e[i]; }
    return n;
  }
  if (fdim == 1) {
    int n = n1 < CCD_MAXDEG ? n1 : CCD_MAXDEG;
    for (int i = 0; i < n; i++) { matvec3(g->rot, g->polynormal + 3 * m1[i], nout[i]); iout[i] = m1[i]; }
    return n;
  }
  return 0;
}
/* :1656-1699 edge directions of a mesh feature: the edge itself, or the edges entering the vertex in each of its polygons */
static int mesh_edge_normals(int dim, const CGeom* g, const real* v1, const real* v2, int v1i, real nout[][3], real endvert[][3]) {
  if (dim == 2) { v3cpy(endvert[0], v2); v3sub(v2, v1, nout[0]); normalize3(nout[0]); return 1; }
  if (dim == 1) {
    const int *m1 = g->polymap + g->polymapadr[v1i];
    int n = g->polymapnum[v1i] < CCD_MAXDEG ? g->polymapnum[v1i] : CCD_MAXDEG;
    for (int i = 0; i < n; i++) {
      const int adr = g->polyvertadr[m1[i]], nvert = g->polyvertnum[m1[i]];
      for (int j = 0; j < nvert; j++)
        if (g->polyvert[adr + j] == v1i) {
          int k = j == 0 ? nvert - 1 : j - 1;
          matvec3(g->rot, g->vert + 3 * g->polyvert[adr + k], endvert[i]);
          for (int c = 0; c < 3; c++) endvert[i][c] += g->pos[c];
          v3sub(endvert[i], v1, nout[i]); normalize3(nout[i]);
        }
    }
    return n;
  }
  return 0;
}
/* :1891-1912 a hull polygon in world coordinates, vertex order reversed */
static int mesh_face(const CGeom* g, int idx, real face[][3]) {
  const int adr = g->polyvertadr[idx], nvert = g->polyvertnum[idx];
  if (nvert > CCD_MAXPOLY) return 0;
  int j = 0;
  for (int i = nvert - 1; i >= 0; i--, j++) {
    matvec3(g->rot, g->vert + 3 * g->polyvert[adr + i], face[j]);
    for (int c = 0; c < 3; c++) face[j][c] += g->pos[c];
  }
  return nvert;
}
/* :2076 multicontact for box / mesh pairs; returns the contact count and overwrites the witness arrays */
static int ccd_multicontact(const Polytope* pt, int epa_face_idx, const real* x1, const real* x2, const CGeom* g1, const CGeom* g2, real w1[4][3], real w2[4][3]) {
  memset(w1, 0, 12 * sizeof(real)); memset(w2, 0, 12 * sizeof(real));
  v3cpy(w1[0], x1); v3cpy(w2[0], x2);
  unsigned f = pt->face[epa_face_idx]; int face[3] = {(int)(f & 0x3FF), (int)((f >> 10) & 0x3FF), (int)((f >> 20) & 0x3FF)};
  int fi1[3], fi2[3]; real fv1[3][3], fv2[3][3];
  int nface1 = feature_dim(pt, face, 0, fi1, fv1), nface2 = feature_dim(pt, face, 1, fi2, fv2);
  real dir[3], dneg[3]; v3sub(x2, x1, dir); for (int i = 0; i < 3; i++) dneg[i] = -dir[i];
  static const int ND = CCD_MAXDEG > 3 ? CCD_MAXDEG : 3;
  real n1[CCD_MAXDEG][3], n2[CCD_MAXDEG][3], endvert[CCD_MAXDEG][3]; int idx1[CCD_MAXDEG], idx2[CCD_MAXDEG];
  (void)ND;
  memset(n1, 0, sizeof n1); memset(n2, 0, sizeof n2); memset(endvert, 0, sizeof endvert); memset(idx1, 0, sizeof idx1); memset(idx2, 0, sizeof idx2);
  int nn1 = g1->type == GEOM_BOX ? box_normals(nface1, fi1, g1->rot, dneg, n1, idx1) : mesh_normals(nface1, fi1, g1, n1, idx1);
  int nn2 = g2->type == GEOM_BOX ? box_normals(nface2, fi2, g2->rot, dir, n2, idx2) : mesh_normals(nface2, fi2, g2, n2, idx2);
  int edge1 = 0, edge2 = 0, ri = 0, rj = 0, found = 0;
  for (int i = 0; i < nn1 && !found; i++) for (int j = 0; j < nn2; j++) if (dot3(n1[i], n2[j]) < -CCD_FACE_TOL) { ri = i; rj = j; found = 1; break; }
  if (!found) {
    if (nface1 < 3 && nface1 <= nface2) {
      nn1 = g1->type == GEOM_BOX ? box_edge_normals(nface1, g1, fv1[0], fv1[1], fi1[0], n1, endvert) : mesh_edge_normals(nface1, g1, fv1[0], fv1[1], fi1[0], n1, endvert);
      for (int i = 0; i < nn2 && !found; i++) for (int j = 0; j < nn1; j++) if (rabs(dot3(n1[j], n2[i])) < CCD_EDGE_TOL) { ri = j; rj = i; found = 1; break; }
      if (!found) return 1;
      edge1 = 1;
    } else if (nface2 < 3) {
      nn2 = g2->type == GEOM_BOX ? box_edge_normals(nface2, g2, fv2[0], fv2[1], fi2[0], n2, endvert) : mesh_edge_normals(nface2, g2, fv2[0], fv2[1], fi2[0], n2, endvert);
      for (int i = 0; i < nn1 && !found; i++) for (int j = 0; j < nn2; j++) if (rabs(dot3(n2[j], n1[i])) < CCD_EDGE_TOL) { ri = j; rj = i; found = 1; break; }
      if (!found) return 1;
      edge2 = 1;
    } else return 1;
  }
  int i = ri, j = rj;
  real face1[CCD_MAXPOLY][3], face2[CCD_MAXPOLY][3];
  if (edge1) { v3cpy(face1[0], pt->vert[2 * face[0]]); v3cpy(face1[1], endvert[i]); nface1 = 2; }
  else { int ind = edge2 ? idx1[j] : idx1[i]; nface1 = g1->type == GEOM_BOX ? box_face(g1, ind, face1) : mesh_face(g1, ind, face1); }
  if (edge2) { v3cpy(face2[0], pt->vert[2 * face[0] + 1]); v3cpy(face2[1], endvert[i]); nface2 = 2; }
  else nface2 = g2->type == GEOM_BOX ? box_face(g2, idx2[j], face2) : mesh_face(g2, idx2[j], face2);
  real dl = len3(dir), ad[3];
  if (edge1) { for (int k = 0; k < 3; k++) ad[k] = -dl * n2[j][k]; return polygon_clip(face2, nface2, face1, nface1, n2[j], ad, w2, w1); } /* faces flipped, flip the witnesses back */
  if (edge2) { for (int k = 0; k < 3; k++) ad[k] = -dl * n1[j][k]; return polygon_clip(face1, nface1, face2, nface2, n1[j], ad, w1, w2); }
  for (int k = 0; k < 3; k++) ad[k] = dl * n2[j][k];
  return polygon_clip(face1, nface1, face2, nface2, n1[i], ad, w1, w2);
}

/* gjk_phase (:2350) + epa_phase (:2421): returns the number of contacts (0 or 1); dist is relative to the margin-inflated shapes */
static int ccd_pair(real tolerance, real cutoff, int gjk_iterations, int epa_iterations, int multi, CGeom g1, CGeom g2, real* dist, real w1[4][3], real w2[4][3], int* ovf) {
  const CGeom o1 = g1, o2 = g2;
  real full1 = 0, full2 = 0, size1 = 0, size2 = 0, *x1 = w1[0], *x2 = w2[0];
  const int disc1 = g1.type == GEOM_BOX || g1.type == GEOM_MESH, disc2 = g2.type == GEOM_BOX || g2.type == GEOM_MESH;
  const int discrete = disc1 && disc2 && g1.margin == 0 && g2.margin == 0; /* :109 _discrete_geoms */
  GjkResult r;
  if (g1.type == GEOM_SPHERE || g1.type == GEOM_CAPSULE) { size1 = g1.size[0]; full1 = size1 + (real)0.5 * g1.margin; g1.margin = 0; g1.size[0] = 0; }
  if (g2.type == GEOM_SPHERE || g2.type == GEOM_CAPSULE) { size2 = g2.size[0]; full2 = size2 + (real)0.5 * g2.margin; g2.margin = 0; g2.size[0] = 0; }
  if (size1 + size2 > 0) {
    cutoff += full1 + full2;
    ccd_gjk(tolerance, gjk_iterations, &g1, &g2, g1.pos, g2.pos, cutoff, discrete, &r);
    if (r.dist > tolerance) {
      v3cpy(x1, r.x1); v3cpy(x2, r.x2);
      if (r.dist == CCD_FLOAT_MAX) { *dist = r.dist; return 1; }
      real n[3]; v3sub(x2, x1, n); normalize3(n); /* :2303 _inflate */
      if (full1 > 0) for (int i = 0; i < 3; i++) x1[i] += full1 * n[i];
      if (full2 > 0) for (int i = 0; i < 3; i++) x2[i] -= full2 * n[i];
      *dist = r.dist - (full1 + full2);
      return 1;
    }
    g1 = o1; g2 = o2; /* margin and size back; the cached support indices stay (:2392-2403) */
    g1.index = r.cindex1; g2.index = r.cindex2;
    cutoff -= full1 + full2;
  }
  ccd_gjk(tolerance, gjk_iterations, &g1, &g2, g1.pos, g2.pos, cutoff, discrete, &r);
  g1.index = r.cindex1; g2.index = r.cindex2;
  if (r.dist > tolerance || r.dim < 2 || r.separated) { *dist = r.dist; v3cpy(x1, r.x1); v3cpy(x2, r.x2); return 1; }
  /* epa_phase */
  int maxvert = 10 + 2 * epa_iterations, maxface = 6 + CCD_MAX_EPAFACES * epa_iterations;
  Polytope pt; memset(&pt, 0, sizeof pt);
  pt.maxvert = maxvert; pt.maxface = maxface;
  pt.vert = (real(*)[3])calloc((size_t)maxvert, 3 * sizeof(real)); pt.vert_index = (int*)calloc((size_t)maxvert, sizeof(int));
  pt.face = (unsigned*)calloc((size_t)maxface, sizeof(unsigned)); pt.face_pr = (real(*)[3])calloc((size_t)maxface, 3 * sizeof(real));
  pt.face_norm2 = (real*)calloc((size_t)maxface, sizeof(real));
  int ncon = 1;
  if (r.dim == 2) { polytope2(&pt, &r, &g1, &g2); if (pt.status == -1) r.dim = 3; }
  else if (r.dim == 4) { polytope4(&pt, &r); if (pt.status == -1) r.dim = 3; }
  if (r.dim == 3) { pt.status = 0; polytope3(&pt, &r, &g1, &g2); }
  if (pt.status) { *dist = r.dist; v3cpy(x1, r.x1); v3cpy(x2, r.x2); }
  else {
    real e1[3], e2[3];
    int fidx = ccd_epa(tolerance, epa_iterations, &pt, &g1, &g2, discrete, dist, e1, e2, ovf);
    if (fidx == -1) { *dist = CCD_FLOAT_MAX; ncon = 0; }
    else {
      v3cpy(x1, e1); v3cpy(x2, e2);
      /* multi-contact: box / mesh pairs without margin (epa_phase :2517-2525, collision_convex.py:875-912); a mesh without polygon
       * data is left at one contact */
      const int mesh_ok = (g1.type != GEOM_MESH || g1.polynum > 0) && (g2.type != GEOM_MESH || g2.polynum > 0);
      if (multi && disc1 && disc2 && mesh_ok && g1.margin == 0 && g2.margin == 0) ncon = ccd_multicontact(&pt, fidx, e1, e2, &g1, &g2, w1, w2);
    }
  }
  free(pt.vert); free(pt.vert_index); free(pt.face); free(pt.face_pr); free(pt.face_norm2);
  return ncon;
}
