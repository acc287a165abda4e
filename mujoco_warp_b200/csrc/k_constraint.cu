// k_constraint.cu -- fused make_constraint: dof-friction rows, joint-limit rows, contact rows (dense Jacobian).
//
// Replaces (reference, /root/reference/mujoco_warp/_src/constraint.py): :61 _zero_constraint_counts, :1765 _friction_dof,
// :1990 _limit_slide_hinge, :2641 _efc_contact_init, :3751 _efc_contact_jac_dense, :4197 _efc_contact_update and the row
// builder :83-152 _efc_row -- ~16 launches with per-row atomics there, one launch here.
//
// One warp owns one world.  Rows are allocated in a fixed order (friction dofs by dof id, limits by joint id via
// ballot/prefix, contacts in the world's contact order), so efc row order is deterministic (the reference's is
// atomics-dependent; its own tests sort before comparing, constraint_test.py:40-59).  Lanes map to dofs: a J row is a
// single coalesced store, J*qvel is a warp-shuffle reduction, and the per-row impedance/reference math runs on the lane
// whose index equals the row's dimension id.
#include "mjb_math.cuh"
#include "mjb_types.cuh"

namespace {

// Per-contact record prepared by ONE lane per contact (phase A: every dependent lookup -- pool fields, geom -> body ->
// root -> subtree_com, invweight -- happens in parallel across contacts), then consumed by the whole warp (phase B: lanes =
// dofs).  The reference does the same work with one thread per contact (_efc_contact_init) and one tile per world
// (_efc_contact_jac_dense), re-reading the pool from global memory in every kernel.
constexpr int CR_FRAME = 0, CR_FRI = 9, CR_OFF1 = 14, CR_OFF2 = 17, CR_POS = 20, CR_INC = 21, CR_INVW = 22, CR_SOLREF = 23,
              CR_SOLREFF = 25, CR_SOLIMP = 27, CR_B1 = 32, CR_B2 = 33, CR_BASE = 34, CR_NDIM = 35, CR_CONDIM = 36, CR_WORDS = 37;
__host__ __device__ inline int con_cap(const DataDev& d) { return 2 * d.nconmax > 32 ? 2 * d.nconmax : 32; }  // = collision's per-world cap

struct ConLayout { int cdof, scom, qvel, rec, total; };
__host__ __device__ inline ConLayout con_layout(const ModelDev& m, const DataDev& d) {
  ConLayout L;
  int o = 0;
  auto take = [&](int n) { int r = o; o += n; return r; };
  L.cdof = take(6 * m.nv); L.scom = take(3 * m.nbody); L.qvel = take(m.nv); L.rec = take(CR_WORDS * 32);  // one 32-contact batch at a time
  L.total = (o + 3) & ~3;
  return L;
}

// constraint.py:83-152
__device__ void efc_row(const ModelDev& m, const DataDev& d, int w, int efcid, float pos_aref, float pos_imp, float invweight,
                        const float* solref, const float* solimp, float margin, float vel, float frictionloss, int type, int id) {
  float timeconst = solref[0];
  const float dampratio = solref[1];
  if (!(m.disableflags & DSBL_REFSAFE)) timeconst = fmaxf(timeconst, 2.0f * m.timestep);
  const float dmin = clampf(solimp[0], MJ_MINIMP, MJ_MAXIMP), dmax = clampf(solimp[1], MJ_MINIMP, MJ_MAXIMP);
  const float width = fmaxf(MJ_MINVAL, solimp[2]), mid = clampf(solimp[3], MJ_MINIMP, MJ_MAXIMP), power = fmaxf(1.0f, solimp[4]);
  const float dmax_sq = dmax * dmax;
  float k = 1.0f / (dmax_sq * timeconst * timeconst * dampratio * dampratio);
  float b = 2.0f / (dmax * timeconst);
  if (solref[0] <= 0.f) k = -solref[0] / dmax_sq;
  if (solref[1] <= 0.f) b = -solref[1] / dmax;
  const float imp_x = fabsf(pos_imp) / width;
  // x^power / c^(power-1) on the lower branch, mirrored on the upper one; power == 2 (MuJoCo default) and 1 avoid powf
  const bool lower = imp_x < mid;
  const float bx = lower ? imp_x : 1.0f - imp_x, bc = lower ? mid : 1.0f - mid;
  float t;
  if (power == 2.0f) t = bx * bx / bc;
  else if (power == 1.0f) t = bx;
  else t = (1.0f / powf(bc, power - 1.0f)) * powf(bx, power);
  const float imp_y = lower ? t : 1.0f - t;
  float imp = clampf(dmin + imp_y * (dmax - dmin), dmin, dmax);
  if (imp_x > 1.0f) imp = dmax;
  const size_t r = (size_t)w * d.njmax + efcid;
  d.efc_D[(size_t)w * d.njmax_pad + efcid] = 1.0f / fmaxf(invweight * (1.0f - imp) / imp, MJ_MINVAL);
  d.efc_vel[r] = vel;
  d.efc_aref[r] = -k * imp * pos_aref - b * vel;
  d.efc_pos[r] = pos_aref + margin;
  d.efc_margin[r] = margin;
  d.efc_frictionloss[r] = frictionloss;
  d.efc_type[r] = type;
  d.efc_id[r] = id;
}

__global__ void __launch_bounds__(MJB_WARPS_PER_BLOCK * 32, 28)
k_constraint(const __grid_constant__ ModelDev m, const __grid_constant__ DataDev d) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x, warp = 0;  // one warp per block: the world index is block-uniform
  const int w = blockIdx.x + d.w0;
  if (w >= d.nworld) return;
  const ConLayout L = con_layout(m, d);
  float* S = smem + warp * L.total;
  float *cdof = S + L.cdof, *scom = S + L.scom, *qvel = S + L.qvel, *rec = S + L.rec;
  const int nv = m.nv, nb = m.nbody, njmax = d.njmax, nvp = d.nv_pad;
  const size_t wb = (size_t)w;
  float* Jw = d.efc_J + wb * (size_t)d.njmax_pad * nvp;   // (nworld, njmax_pad, nv_pad)

  if (m.disableflags & DSBL_CONSTRAINT) {
    if (lane == 0) { d.ne[w] = 0; d.nf[w] = 0; d.nl[w] = 0; d.nefc[w] = 0; }
    return;
  }
  warp_copy(cdof, d.cdof + wb * 6 * nv, 6 * nv, lane);
  warp_copy(scom, d.subtree_com + wb * 3 * nb, 3 * nb, lane);
  warp_copy(qvel, d.qvel + wb * nv, nv, lane);
  __syncwarp();

  int nefc = 0, nf = 0, nl = 0;

  // ---- dof friction loss rows (always present when frictionloss > 0)
  if (!(m.disableflags & DSBL_FRICTIONLOSS)) {
    nf = m.nfricdof;
    for (int i = 0; i < nf; i++) {
      const int efcid = nefc + i, dof = m.dof_fricloss_adr[i];
      if (efcid >= njmax) break;
#pragma unroll 1
      for (int c = lane; c < nvp; c += 32) Jw[(size_t)efcid * nvp + c] = c == dof ? 1.0f : 0.f;
      if (lane == 0)
        efc_row(m, d, w, efcid, 0.f, 0.f, m.dof_invweight0[dof], m.dof_solref + 2 * dof, m.dof_solimp + 5 * dof, 0.f, qvel[dof],
                m.dof_frictionloss[dof], CNSTR_FRICTION_DOF, dof);
    }
    nefc += nf;
  }

  // ---- joint limits (slide / hinge)
  if (!(m.disableflags & DSBL_LIMIT)) {
#pragma unroll 1
    for (int l0 = 0; l0 < m.nlimit; l0 += 32) {
      const int li = l0 + lane;
      bool active = false;
      int j = 0, dofadr = 0;
      float pos = 0.f, Jv = 0.f, margin = 0.f;
      if (li < m.nlimit) {
        j = m.jnt_limited_adr[li];
        const float q = d.qpos[wb * m.nq + m.jnt_qposadr[j]];
        margin = m.jnt_margin[j];
        const float dist_min = q - m.jnt_range[2 * j], dist_max = m.jnt_range[2 * j + 1] - q;
        pos = fminf(dist_min, dist_max) - margin;
        active = pos < 0.f;
        Jv = dist_min < dist_max ? 1.0f : -1.0f;
        dofadr = m.jnt_dofadr[j];
      }
      const unsigned bal = __ballot_sync(FULL_MASK, active);
      const int efcid = nefc + __popc(bal & ((1u << lane) - 1u));
      if (active && efcid < njmax)
        efc_row(m, d, w, efcid, pos, pos, m.dof_invweight0[dofadr], m.jnt_solref + 2 * j, m.jnt_solimp + 5 * j, margin,
                Jv * qvel[dofadr], 0.f, CNSTR_LIMIT_JOINT, j);
      unsigned rem = bal;
      while (rem) {  // whole warp writes each active row's J
        const int src = __ffs(rem) - 1;
        rem &= rem - 1;
        const int r = __shfl_sync(FULL_MASK, efcid, src), dcol = __shfl_sync(FULL_MASK, dofadr, src);
        const float jv = __shfl_sync(FULL_MASK, Jv, src);
        if (r < njmax)
#pragma unroll 1
          for (int c = lane; c < nvp; c += 32) Jw[(size_t)r * nvp + c] = c == dcol ? jv : 0.f;
      }
      const int n = __popc(bal);
      nefc += n; nl += n;
    }
  }

  // ---- contacts
  if (!(m.disableflags & DSBL_CONTACT)) {
    const int cbase = d.world_conadr[w], ncon = min(d.world_ncon[w], con_cap(d)), np = m.nmaxpyramid;
    const bool elliptic = m.cone == CONE_ELLIPTIC;
#pragma unroll 1
    for (int c0 = 0; c0 < ncon; c0 += 32) {
      // ---- phase A: lane = contact (batch of 32)
      const int c = c0 + lane, cid = cbase + c;
      int ndim = 0, condim = 0;
      float includemargin = 0.f, pos = 0.f;
      if (c < ncon) {
        includemargin = d.contact_includemargin[cid];
        pos = d.contact_dist[cid] - includemargin;
        condim = d.contact_dim[cid];
        if (pos < 0.f) ndim = elliptic ? condim : (condim == 1 ? 1 : 2 * (condim - 1));
      }
      const int base = nefc + warp_excl_scan(ndim, lane);
      nefc += warp_sum_i(ndim);
      if (c < ncon) {
        float* r = rec + CR_WORDS * lane;
        r[CR_NDIM] = __int_as_float(ndim);
        if (ndim > 0) {
          const int g1 = d.contact_geom[2 * cid], g2 = d.contact_geom[2 * cid + 1], b1 = m.geom_bodyid[g1], b2 = m.geom_bodyid[g2];
          const v3 cpos = ld3(d.contact_pos + 3 * cid);
          for (int k = 0; k < 9; k++) r[CR_FRAME + k] = d.contact_frame[9 * (size_t)cid + k];
          for (int k = 0; k < 5; k++) { r[CR_FRI + k] = d.contact_friction[5 * cid + k]; r[CR_SOLIMP + k] = d.contact_solimp[5 * cid + k]; }
          st3(r + CR_OFF1, cpos - ld3(scom + 3 * m.body_rootid[b1]));
          st3(r + CR_OFF2, cpos - ld3(scom + 3 * m.body_rootid[b2]));
          r[CR_POS] = pos; r[CR_INC] = includemargin;
          r[CR_INVW] = m.body_invweight0[2 * b1] + m.body_invweight0[2 * b2];
          r[CR_SOLREF] = d.contact_solref[2 * cid]; r[CR_SOLREF + 1] = d.contact_solref[2 * cid + 1];
          r[CR_SOLREFF] = d.contact_solreffriction[2 * cid]; r[CR_SOLREFF + 1] = d.contact_solreffriction[2 * cid + 1];
          r[CR_B1] = __int_as_float(b1); r[CR_B2] = __int_as_float(b2);
          r[CR_BASE] = __int_as_float(base); r[CR_CONDIM] = __int_as_float(condim);
          for (int k = 0; k < ndim; k++) d.contact_efc_address[np * cid + k] = base + k < njmax ? base + k : -1;
        }
      }
      __syncwarp();
      // ---- phase B: lanes = dofs, the batch's contacts one after the other, every operand already in shared memory
      const int nb_ = min(32, ncon - c0);
#pragma unroll 1
      for (int cb = 0; cb < nb_; cb++) {
      const float* r = rec + CR_WORDS * cb;
      const int ndim = __float_as_int(r[CR_NDIM]);
      if (ndim == 0) continue;
      const int cid = cbase + c0 + cb, base = __float_as_int(r[CR_BASE]), condim = __float_as_int(r[CR_CONDIM]);
      const int b1 = __float_as_int(r[CR_B1]), b2 = __float_as_int(r[CR_B2]);
      const float pos = r[CR_POS], includemargin = r[CR_INC];
      float frame[9], fri[5];
      for (int k = 0; k < 9; k++) frame[k] = r[CR_FRAME + k];
      for (int k = 0; k < 5; k++) fri[k] = r[CR_FRI + k];
      const v3 off1 = ld3(r + CR_OFF1), off2 = ld3(r + CR_OFF2);
      float velp[10];
#pragma unroll
      for (int k = 0; k < 10; k++) velp[k] = 0.f;
#pragma unroll 1
      for (int dd = lane; dd < nvp; dd += 32) {
        v3 jpd = mk3(0.f, 0.f, 0.f), jrd = mk3(0.f, 0.f, 0.f);
        float qv = 0.f;
        if (dd < nv) {
          const v3 ang = ld3(cdof + 6 * dd), lin = ld3(cdof + 6 * dd + 3);
          const int a1 = m.body_isdofancestor[b1 * nv + dd], a2 = m.body_isdofancestor[b2 * nv + dd];
          if (a2) { jpd = lin + cross(ang, off2); jrd = ang; }
          if (a1) { jpd = jpd - (lin + cross(ang, off1)); jrd = jrd - ang; }
          qv = qvel[dd];
        }
        const float p0 = dot(jpd, ld3(frame)), p1 = dot(jpd, ld3(frame + 3)), p2 = dot(jpd, ld3(frame + 6));
        const float r0 = dot(jrd, ld3(frame)), r1 = dot(jrd, ld3(frame + 3)), r2 = dot(jrd, ld3(frame + 6));
#pragma unroll
        for (int dim = 0; dim < 10; dim++) {
          if (dim < ndim) {
            float J;
            if (elliptic) {
              J = dim == 0 ? p0 : dim == 1 ? p1 : dim == 2 ? p2 : dim == 3 ? r0 : dim == 4 ? r1 : r2;
            } else {
              J = p0;
              if (condim > 1) {
                const int dimid2 = dim / 2 + 1;
                const float frii = fri[dimid2 - 1] * ((dim & 1) ? -1.0f : 1.0f);
                const float comp = dimid2 == 1 ? p1 : dimid2 == 2 ? p2 : dimid2 == 3 ? r0 : dimid2 == 4 ? r1 : r2;
                J += comp * frii;
              }
            }
            if (base + dim < njmax) Jw[(size_t)(base + dim) * nvp + dd] = J;
            velp[dim] += J * qv;
          }
        }
      }
      float myvel = 0.f;
#pragma unroll
      for (int dim = 0; dim < 10; dim++) {
        if (dim < ndim) { const float s = warp_sum(velp[dim]); if (lane == dim) myvel = s; }
      }
      if (lane < ndim) {  // constraint.py:4197-4343
        const int dim = lane, efcid = base + dim;
        if (efcid < njmax) {
          float invweight = r[CR_INVW];
          float pos_aref = pos;
          float ref[2] = {r[CR_SOLREF], r[CR_SOLREF + 1]};
          float imp5[5];
          for (int k = 0; k < 5; k++) imp5[k] = r[CR_SOLIMP + k];
          if (elliptic) {
            if (dim > 0) {
              const float s0 = r[CR_SOLREFF], s1 = r[CR_SOLREFF + 1];
              if (s0 != 0.f || s1 != 0.f) { ref[0] = s0; ref[1] = s1; }
              invweight = invweight * m.impratio_invsqrt * m.impratio_invsqrt;
              if (dim > 1) invweight *= fri[0] * fri[0] / (fri[dim - 1] * fri[dim - 1]);
              pos_aref = 0.f;
            }
          } else if (condim > 1) {
            const float f0 = fri[0];
            invweight = invweight + f0 * f0 * invweight;
            invweight = invweight * 2.0f * f0 * f0 * m.impratio_invsqrt * m.impratio_invsqrt;
          }
          const int type = condim == 1 ? CNSTR_CONTACT_FRICTIONLESS : (elliptic ? CNSTR_CONTACT_ELLIPTIC : CNSTR_CONTACT_PYRAMIDAL);
          efc_row(m, d, w, efcid, pos_aref, pos, invweight, ref, imp5, includemargin, myvel, 0.f, type, cid);
        }
      }
      }
      __syncwarp();
    }
  }
  if (lane == 0) { d.ne[w] = 0; d.nf[w] = nf; d.nl[w] = nl; d.nefc[w] = nefc; }
}

}  // namespace

size_t smem_constraint(const ModelDev& m, const DataDev& d) { return (size_t)con_layout(m, d).total * sizeof(float) * MJB_WARPS_PER_BLOCK; }

cudaError_t launch_constraint(const ModelDev& m, const DataDev& d, cudaStream_t s) {
  const size_t smem = smem_constraint(m, d);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_constraint, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const int grid = d.wn;
  k_constraint<<<grid, MJB_WARPS_PER_BLOCK * 32, smem, s>>>(m, d);
  return cudaGetLastError();
}
