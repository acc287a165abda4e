// k_position.cu -- fused position stage: kinematics -> com_pos -> camlight -> crb (+M) -> transmission.
//
// Replaces (reference, /root/reference/mujoco_warp/_src/): smooth.py:46-226 (5 kinematics kernels),
// :686-855 (com_pos: 11 launches with float atomics per tree level), :858-1027 (camlight), :1029-1098 (crb: 10
// launches), :2288-2396 (_transmission, joint transmission).  One warp owns one world: the body tree lives in the
// warp's shared-memory slice, tree passes run level by level with __syncwarp() instead of one launch per level,
// parents gather their children in fixed order (deterministic; no float atomics), and every Data field is written
// once with coalesced row stores.  `mask` selects sub-stages so each public stage function stays individually callable;
// inputs a skipped sub-stage would have produced are re-loaded from Data.
#include "mjb_math.cuh"
#include "mjb_types.cuh"

namespace {

struct PosLayout {
  int qpos, xpos, xquat, xmat, xipos, ximat, xanchor, xaxis, gxpos, gxmat, scom, cinert, crb, cdof, buf, M, total;
};

__host__ __device__ inline PosLayout pos_layout(const ModelDev& m) {
  PosLayout L;
  int o = 0;
  auto take = [&](int n) { int r = o; o += n; return r; };
  L.qpos = take(m.nq);
  L.xpos = take(3 * m.nbody); L.xquat = take(4 * m.nbody); L.xmat = take(9 * m.nbody);
  L.xipos = take(3 * m.nbody); L.ximat = take(9 * m.nbody);
  L.xanchor = take(3 * m.njnt); L.xaxis = take(3 * m.njnt);
  L.gxpos = take(3 * m.ngeom); L.gxmat = take(9 * m.ngeom);
  L.scom = take(3 * m.nbody); L.cinert = take(10 * m.nbody); L.crb = take(10 * m.nbody);
  L.cdof = take(6 * m.nv); L.M = take(m.nC);
  // crb*cdof scratch (6 nv) reuses the geom_xmat staging area, which is dead once kinematics has been written out
  L.buf = L.gxmat;
  if (6 * m.nv > 9 * m.ngeom) L.buf = take(6 * m.nv);
  L.total = (o + 3) & ~3;
  return L;
}

__global__ void __launch_bounds__(MJB_WARPS_PER_BLOCK * 32)
k_position(const __grid_constant__ ModelDev m, const __grid_constant__ DataDev d, int mask) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x, warp = 0;  // one warp per block: the world index is block-uniform
  const int w = blockIdx.x + d.w0;
  if (w >= d.nworld) return;
  const PosLayout L = pos_layout(m);
  float* S = smem + warp * L.total;
  float *qpos = S + L.qpos, *xpos = S + L.xpos, *xquat = S + L.xquat, *xmat = S + L.xmat, *xipos = S + L.xipos,
        *ximat = S + L.ximat, *xanchor = S + L.xanchor, *xaxis = S + L.xaxis, *gxpos = S + L.gxpos, *gxmat = S + L.gxmat,
        *scom = S + L.scom, *cinert = S + L.cinert, *crb = S + L.crb, *cdof = S + L.cdof, *buf = S + L.buf, *Ms = S + L.M;
  const int nb = m.nbody, nj = m.njnt, ng = m.ngeom, nv = m.nv;
  const size_t wb = (size_t)w;

  warp_copy(qpos, d.qpos + wb * m.nq, m.nq, lane);

  // ------------------------------------------------------------------ kinematics
  if (mask & STG_KINEMATICS) {
    if (lane == 0) { xpos[0] = xpos[1] = xpos[2] = 0.f; xquat[0] = 1.f; xquat[1] = xquat[2] = xquat[3] = 0.f; }
    __syncwarp();
#pragma unroll 1
    for (int l = 1; l < m.nlevel; l++) {
#pragma unroll 1
      for (int i = m.level_adr[l] + lane; i < m.level_adr[l + 1]; i += 32) {
        const int b = m.level_body[i], pid = m.body_parentid[b], jntadr = m.body_jntadr[b], jntnum = m.body_jntnum[b];
        if (jntnum == 1 && m.jnt_type[jntadr] == JNT_FREE) {
          const int qa = m.jnt_qposadr[jntadr];
          v3 p = ld3(qpos + qa);
          q4 q = qnormalize(ldq(qpos + qa + 3));
          st3(xpos + 3 * b, p); stq(xquat + 4 * b, q);
          st3(xanchor + 3 * jntadr, p); st3(xaxis + 3 * jntadr, ld3(m.jnt_axis + 3 * jntadr));
          continue;
        }
        q4 pq = ldq(xquat + 4 * pid);
        // mocap bodies take their pose from Data.mocap_pos / mocap_quat (smooth.py:104-110)
        const int mc = m.nmocap > 0 ? m.body_mocapid[b] : -1;
        const v3 bpos = mc >= 0 ? ld3(d.mocap_pos + (wb * m.nmocap + mc) * 3) : ld3(m.body_pos + 3 * b);
        const q4 bquat = mc >= 0 ? ldq(d.mocap_quat + (wb * m.nmocap + mc) * 4) : ldq(m.body_quat + 4 * b);
        v3 pos = qrot(pq, bpos) + ld3(xpos + 3 * pid);
        q4 quat = qmul(pq, bquat);
#pragma unroll 1
        for (int j = jntadr; j < jntadr + jntnum; j++) {
          const int qa = m.jnt_qposadr[j], t = m.jnt_type[j];
          v3 jpos = ld3(m.jnt_pos + 3 * j), jax = ld3(m.jnt_axis + 3 * j);
          v3 anchor = qrot(quat, jpos) + pos, axis = qrot(quat, jax);
          if (t == JNT_BALL) {
            quat = qmul(quat, qnormalize(ldq(qpos + qa)));
            pos = anchor - qrot(quat, jpos);
          } else if (t == JNT_SLIDE) {
            pos = pos + axis * (qpos[qa] - m.qpos0[qa]);
          } else if (t == JNT_HINGE) {
            quat = qmul(quat, axis_angle_quat(jax, qpos[qa] - m.qpos0[qa]));
            pos = anchor - qrot(quat, jpos);
          }
          st3(xanchor + 3 * j, anchor); st3(xaxis + 3 * j, axis);
        }
        st3(xpos + 3 * b, pos); stq(xquat + 4 * b, qnormalize(quat));
      }
      __syncwarp();
    }
#pragma unroll 1
    for (int b = lane; b < nb; b += 32) {
      q4 q = ldq(xquat + 4 * b);
      quat_to_mat(q, xmat + 9 * b);
      st3(xipos + 3 * b, ld3(xpos + 3 * b) + qrot(q, ld3(m.body_ipos + 3 * b)));
      quat_to_mat(qmul(q, ldq(m.body_iquat + 4 * b)), ximat + 9 * b);
    }
#pragma unroll 1
    for (int g = lane; g < ng; g += 32) {
      const int b = m.geom_bodyid[g];
      if (m.body_weldid[b] == 0 && (m.nmocap == 0 || m.body_mocapid[m.body_rootid[b]] == -1)) {  // static geom: keeps the pose computed at make_data (smooth.py:197-200)
        for (int k = 0; k < 3; k++) gxpos[3 * g + k] = d.geom_xpos[(wb * ng + g) * 3 + k];
        for (int k = 0; k < 9; k++) gxmat[9 * g + k] = d.geom_xmat[(wb * ng + g) * 9 + k];
      } else {
        q4 q = ldq(xquat + 4 * b);
        st3(gxpos + 3 * g, ld3(xpos + 3 * b) + qrot(q, ld3(m.geom_pos + 3 * g)));
        quat_to_mat(qmul(q, ldq(m.geom_quat + 4 * g)), gxmat + 9 * g);
      }
    }
#pragma unroll 1
    for (int s = lane; s < m.nsite; s += 32) {
      const int b = m.site_bodyid[s];
      q4 q = ldq(xquat + 4 * b);
      float mat[9];
      st3(d.site_xpos + (wb * m.nsite + s) * 3, ld3(xpos + 3 * b) + qrot(q, ld3(m.site_pos + 3 * s)));
      quat_to_mat(qmul(q, ldq(m.site_quat + 4 * s)), mat);
      for (int k = 0; k < 9; k++) d.site_xmat[(wb * m.nsite + s) * 9 + k] = mat[k];
    }
    __syncwarp();
    warp_copy(d.xpos + wb * 3 * nb, xpos, 3 * nb, lane);
    warp_copy(d.xquat + wb * 4 * nb, xquat, 4 * nb, lane);
    warp_copy(d.xmat + wb * 9 * nb, xmat, 9 * nb, lane);
    warp_copy(d.xipos + wb * 3 * nb, xipos, 3 * nb, lane);
    warp_copy(d.ximat + wb * 9 * nb, ximat, 9 * nb, lane);
    warp_copy(d.xanchor + wb * 3 * nj, xanchor, 3 * nj, lane);
    warp_copy(d.xaxis + wb * 3 * nj, xaxis, 3 * nj, lane);
    warp_copy(d.geom_xpos + wb * 3 * ng, gxpos, 3 * ng, lane);
    warp_copy(d.geom_xmat + wb * 9 * ng, gxmat, 9 * ng, lane);
  } else if (mask & (STG_COM_POS | STG_CAMLIGHT)) {
    warp_copy(xpos, d.xpos + wb * 3 * nb, 3 * nb, lane);
    warp_copy(xquat, d.xquat + wb * 4 * nb, 4 * nb, lane);
    warp_copy(xmat, d.xmat + wb * 9 * nb, 9 * nb, lane);
    warp_copy(xipos, d.xipos + wb * 3 * nb, 3 * nb, lane);
    warp_copy(ximat, d.ximat + wb * 9 * nb, 9 * nb, lane);
    warp_copy(xanchor, d.xanchor + wb * 3 * nj, 3 * nj, lane);
    warp_copy(xaxis, d.xaxis + wb * 3 * nj, 3 * nj, lane);
  }
  __syncwarp();

  // ------------------------------------------------------------------ com_pos
  if (mask & STG_COM_POS) {
#pragma unroll 1
    for (int b = lane; b < nb; b += 32) st3(scom + 3 * b, ld3(xipos + 3 * b) * m.body_mass[b]);
    __syncwarp();
#pragma unroll 1
    for (int l = m.nlevel - 2; l >= 0; l--) {
#pragma unroll 1
      for (int i = m.level_adr[l] + lane; i < m.level_adr[l + 1]; i += 32) {
        const int b = m.level_body[i];
        v3 acc = ld3(scom + 3 * b);
        for (int c = m.body_childadr[b]; c < m.body_childadr[b + 1]; c++) acc = acc + ld3(scom + 3 * m.body_childid[c]);
        st3(scom + 3 * b, acc);
      }
      __syncwarp();
    }
#pragma unroll 1
    for (int b = lane; b < nb; b += 32) {
      const float ms = m.body_subtreemass[b];
      if (ms != 0.f) st3(scom + 3 * b, ld3(scom + 3 * b) * (1.0f / ms));
    }
    __syncwarp();
#pragma unroll 1
    for (int b = lane; b < nb; b += 32) {  // cinert (smooth.py:733)
      const float* mat = ximat + 9 * b;
      const v3 inert = ld3(m.body_inertia + 3 * b);
      const float mass = m.body_mass[b];
      const v3 dif = ld3(xipos + 3 * b) - ld3(scom + 3 * m.body_rootid[b]);
      float* r = cinert + 10 * b;
      // mat * diag(inert) * mat^T, symmetric
      const float a0 = mat[0] * inert.x, a1 = mat[1] * inert.y, a2 = mat[2] * inert.z;
      const float b0 = mat[3] * inert.x, b1 = mat[4] * inert.y, b2 = mat[5] * inert.z;
      const float c0 = mat[6] * inert.x, c1 = mat[7] * inert.y, c2 = mat[8] * inert.z;
      r[0] = a0 * mat[0] + a1 * mat[1] + a2 * mat[2] + mass * (dif.y * dif.y + dif.z * dif.z);
      r[1] = b0 * mat[3] + b1 * mat[4] + b2 * mat[5] + mass * (dif.x * dif.x + dif.z * dif.z);
      r[2] = c0 * mat[6] + c1 * mat[7] + c2 * mat[8] + mass * (dif.x * dif.x + dif.y * dif.y);
      r[3] = a0 * mat[3] + a1 * mat[4] + a2 * mat[5] - mass * dif.x * dif.y;
      r[4] = a0 * mat[6] + a1 * mat[7] + a2 * mat[8] - mass * dif.x * dif.z;
      r[5] = b0 * mat[6] + b1 * mat[7] + b2 * mat[8] - mass * dif.y * dif.z;
      r[6] = mass * dif.x; r[7] = mass * dif.y; r[8] = mass * dif.z; r[9] = mass;
    }
#pragma unroll 1
    for (int j = lane; j < nj; j += 32) {  // cdof (smooth.py:779)
      const int b = m.jnt_bodyid[j], t = m.jnt_type[j];
      int dof = m.jnt_dofadr[j];
      const v3 offset = ld3(scom + 3 * m.body_rootid[b]) - ld3(xanchor + 3 * j);
      const float* xm = xmat + 9 * b;
      if (t == JNT_FREE || t == JNT_BALL) {
        if (t == JNT_FREE) {
          for (int k = 0; k < 18; k++) cdof[6 * dof + k] = 0.f;
          cdof[6 * dof + 3] = 1.f; cdof[6 * (dof + 1) + 4] = 1.f; cdof[6 * (dof + 2) + 5] = 1.f;
          dof += 3;
        }
        for (int k = 0; k < 3; k++) {
          v3 col = matcol(xm, k);
          st3(cdof + 6 * (dof + k), col); st3(cdof + 6 * (dof + k) + 3, cross(col, offset));
        }
      } else if (t == JNT_SLIDE) {
        st3(cdof + 6 * dof, mk3(0.f, 0.f, 0.f)); st3(cdof + 6 * dof + 3, ld3(xaxis + 3 * j));
      } else {
        v3 ax = ld3(xaxis + 3 * j);
        st3(cdof + 6 * dof, ax); st3(cdof + 6 * dof + 3, cross(ax, offset));
      }
    }
    __syncwarp();
    warp_copy(d.subtree_com + wb * 3 * nb, scom, 3 * nb, lane);
    warp_copy(d.cinert + wb * 10 * nb, cinert, 10 * nb, lane);
    warp_copy(d.cdof + wb * 6 * nv, cdof, 6 * nv, lane);
  } else if (mask & (STG_CAMLIGHT | STG_CRB)) {
    warp_copy(scom, d.subtree_com + wb * 3 * nb, 3 * nb, lane);
    warp_copy(cinert, d.cinert + wb * 10 * nb, 10 * nb, lane);
    warp_copy(cdof, d.cdof + wb * 6 * nv, 6 * nv, lane);
  }
  __syncwarp();

  // ------------------------------------------------------------------ camlight (smooth.py:858-1027)
  if (mask & STG_CAMLIGHT) {
#pragma unroll 1
    for (int c = lane; c < m.ncam; c += 32) {
      const int mode = m.cam_mode[c], b = m.cam_bodyid[c], tb = m.cam_targetbodyid[c];
      const bool is_target = mode == CAM_TARGETBODY || mode == CAM_TARGETBODYCOM;
      v3 p; float mat[9];
      if (mode == CAM_TRACK) {
        for (int k = 0; k < 9; k++) mat[k] = m.cam_mat0[9 * c + k];
        p = ld3(xpos + 3 * b) + ld3(m.cam_pos0 + 3 * c);
      } else if (mode == CAM_TRACKCOM) {
        for (int k = 0; k < 9; k++) mat[k] = m.cam_mat0[9 * c + k];
        p = ld3(scom + 3 * b) + ld3(m.cam_poscom0 + 3 * c);
      } else if (is_target && tb >= 0) {
        q4 q = ldq(xquat + 4 * b);
        p = ld3(xpos + 3 * b) + qrot(q, ld3(m.cam_pos + 3 * c));
        v3 tp = mode == CAM_TARGETBODYCOM ? ld3(scom + 3 * tb) : ld3(xpos + 3 * tb);
        v3 m3 = normalize(p - tp), m1 = normalize(cross(mk3(0.f, 0.f, 1.f), m3)), m2 = normalize(cross(m3, m1));
        mat[0] = m1.x; mat[1] = m2.x; mat[2] = m3.x; mat[3] = m1.y; mat[4] = m2.y; mat[5] = m3.y; mat[6] = m1.z; mat[7] = m2.z; mat[8] = m3.z;
      } else {
        q4 q = ldq(xquat + 4 * b);
        p = ld3(xpos + 3 * b) + qrot(q, ld3(m.cam_pos + 3 * c));
        quat_to_mat(qmul(q, ldq(m.cam_quat + 4 * c)), mat);
      }
      st3(d.cam_xpos + (wb * m.ncam + c) * 3, p);
      for (int k = 0; k < 9; k++) d.cam_xmat[(wb * m.ncam + c) * 9 + k] = mat[k];
    }
#pragma unroll 1
    for (int l = lane; l < m.nlight; l += 32) {
      const int mode = m.light_mode[l], b = m.light_bodyid[l], tb = m.light_targetbodyid[l];
      const bool is_target = mode == CAM_TARGETBODY || mode == CAM_TARGETBODYCOM;
      v3 p, dir;
      bool norm = true;
      q4 q = ldq(xquat + 4 * b);
      if (is_target && tb < 0) {
        p = ld3(xpos + 3 * b) + qrot(q, ld3(m.light_pos + 3 * l)); dir = qrot(q, ld3(m.light_dir + 3 * l)); norm = false;
      } else if (mode == CAM_TRACK) {
        dir = ld3(m.light_dir0 + 3 * l); p = ld3(xpos + 3 * b) + ld3(m.light_pos0 + 3 * l);
      } else if (mode == CAM_TRACKCOM) {
        dir = ld3(m.light_dir0 + 3 * l); p = ld3(scom + 3 * b) + ld3(m.light_poscom0 + 3 * l);
      } else if (is_target) {
        p = ld3(xpos + 3 * b) + qrot(q, ld3(m.light_pos + 3 * l));
        v3 tp = mode == CAM_TARGETBODYCOM ? ld3(scom + 3 * tb) : ld3(xpos + 3 * tb);
        dir = tp - p;
      } else {
        p = ld3(xpos + 3 * b) + qrot(q, ld3(m.light_pos + 3 * l)); dir = qrot(q, ld3(m.light_dir + 3 * l));
      }
      if (norm) dir = normalize(dir);
      st3(d.light_xpos + (wb * m.nlight + l) * 3, p);
      st3(d.light_xdir + (wb * m.nlight + l) * 3, dir);
    }
  }

  // ------------------------------------------------------------------ crb + M (smooth.py:1029-1098)
  if (mask & STG_CRB) {
#pragma unroll 1
    for (int i = lane; i < 10 * nb; i += 32) crb[i] = cinert[i];
    __syncwarp();
#pragma unroll 1
    for (int l = m.nlevel - 2; l >= 1; l--) {
#pragma unroll 1
      for (int i = m.level_adr[l] + lane; i < m.level_adr[l + 1]; i += 32) {
        const int b = m.level_body[i];
        float acc[10];
#pragma unroll
        for (int k = 0; k < 10; k++) acc[k] = crb[10 * b + k];
#pragma unroll 1
        for (int c = m.body_childadr[b]; c < m.body_childadr[b + 1]; c++) {
          const float* cc = crb + 10 * m.body_childid[c];
#pragma unroll
          for (int k = 0; k < 10; k++) acc[k] += cc[k];
        }
#pragma unroll
        for (int k = 0; k < 10; k++) crb[10 * b + k] = acc[k];
      }
      __syncwarp();
    }
#pragma unroll 1
    for (int dd = lane; dd < nv; dd += 32) inert_vec(crb + 10 * m.dof_bodyid[dd], cdof + 6 * dd, buf + 6 * dd);
    __syncwarp();
#pragma unroll 1
    for (int e = lane; e < m.nC; e += 32) {
      const int i = m.M_entry_row[e], j = m.M_colind[e];
      float v = dot6(cdof + 6 * j, buf + 6 * i);
      if (i == j) v += m.dof_armature[i];
      Ms[e] = v;
    }
    __syncwarp();
    warp_copy(d.crb + wb * 10 * nb, crb, 10 * nb, lane);
    warp_copy(d.M + wb * m.nC, Ms, m.nC, lane);
  }

  // ------------------------------------------------------------------ transmission (joint transmission; smooth.py:2288-2396)
  if (mask & STG_TRANSMISSION) {
#pragma unroll 1
    for (int a = lane; a < m.nu; a += 32) {
      const int j = m.actuator_trnid[2 * a], t = m.jnt_type[j], adr = m.moment_rowadr0[a], nnz = m.moment_rownnz0[a];
      const float* gear = m.actuator_gear + 6 * a;
      d.actuator_length[wb * m.nu + a] = (t == JNT_SLIDE || t == JNT_HINGE) ? qpos[m.jnt_qposadr[j]] * gear[0] : 0.f;
      d.moment_rownnz[wb * m.nu + a] = nnz;
      d.moment_rowadr[wb * m.nu + a] = adr;
#pragma unroll 1
      for (int k = 0; k < nnz; k++) {
        d.moment_colind[wb * m.nJmom + adr + k] = m.moment_colind0[adr + k];
        d.actuator_moment[wb * m.nJmom + adr + k] = gear[k];
      }
    }
  }
}

}  // namespace

size_t smem_position(const ModelDev& m) { return (size_t)pos_layout(m).total * sizeof(float) * MJB_WARPS_PER_BLOCK; }

cudaError_t launch_position(const ModelDev& m, const DataDev& d, int mask, cudaStream_t s) {
  const size_t smem = smem_position(m);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_position, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const int grid = d.wn;
  k_position<<<grid, MJB_WARPS_PER_BLOCK * 32, smem, s>>>(m, d, mask);
  return cudaGetLastError();
}
