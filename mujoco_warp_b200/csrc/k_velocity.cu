// k_velocity.cu -- fused velocity / actuation / acceleration stage.
//
// Replaces (reference, /root/reference/mujoco_warp/_src/): forward.py:680 _actuator_velocity; smooth.py:2179-2285 com_vel;
// passive.py:73-206,631-667 (joint springs/dampers, passive sum); smooth.py:1353-1515 rne (incl. 7 per-level atomic
// launches of _cfrc_backward); forward.py:756-1149 fwd_actuation (stateless FIXED/AFFINE gain, NONE/AFFINE bias, joint
// transmission); forward.py:1255-1324 fwd_acceleration with support.py:259-324 xfrc_accumulate and the per-tree dense
// Cholesky factor+solve of M (smooth.py:3227-3265) -- about 30 launches there, one here.
//
// One warp owns one world; tree passes are level-synchronous in shared memory, children are gathered by the parent in a
// fixed order (no float atomics => bit-reproducible), the inertia block is factored by the warp in shared memory.
#include "mjb_chol.cuh"
#include "mjb_math.cuh"
#include "mjb_types.cuh"

namespace {

struct VelLayout { int qvel, cdof, cinert, cvel, cdofdot, cacc, cfrc, qf, A, x, af, total; };
__host__ __device__ inline int chol_ld(int n) { return (n | 1); }  // odd leading dimension
__host__ __device__ inline VelLayout vel_layout(const ModelDev& m) {
  VelLayout L;
  int o = 0;
  auto take = [&](int n) { int r = o; o += n; return r; };
  L.qvel = take(m.nv); L.cdof = take(6 * m.nv); L.cinert = take(10 * m.nbody); L.cvel = take(6 * m.nbody);
  L.cdofdot = take(6 * m.nv); L.cacc = take(6 * m.nbody); L.cfrc = take(6 * m.nbody);
  L.qf = take(4 * m.nv);  // passive, bias, actuator, smooth
  L.A = take(m.maxtree * chol_ld(m.maxtree)); L.x = take(m.maxtree); L.af = take(m.nu);
  L.total = (o + 3) & ~3;
  return L;
}

// 28 resident one-warp blocks per SM make 8192 worlds exactly two full waves on 148 SMs (caps registers at 72)
// PEXT = the model uses gravity compensation or free / ball joint springs (kept out of the plain instantiation)
template <bool PEXT>
__global__ void __launch_bounds__(MJB_WARPS_PER_BLOCK * 32, 28)
k_velocity(const __grid_constant__ ModelDev m, const __grid_constant__ DataDev d, int mask) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x, warp = 0;  // one warp per block: the world index is block-uniform
  const int w = blockIdx.x + d.w0;
  if (w >= d.nworld) return;
  const VelLayout L = vel_layout(m);
  float* S = smem + warp * L.total;
  float *qvel = S + L.qvel, *cdof = S + L.cdof, *cinert = S + L.cinert, *cvel = S + L.cvel, *cdofdot = S + L.cdofdot,
        *cacc = S + L.cacc, *cfrc = S + L.cfrc, *A = S + L.A, *x = S + L.x, *aforce = S + L.af;
  float *q_passive = S + L.qf, *q_bias = q_passive + m.nv, *q_act = q_bias + m.nv, *q_smooth = q_act + m.nv;
  const int nv = m.nv, nb = m.nbody, nu = m.nu;
  const size_t wb = (size_t)w;

  warp_copy(qvel, d.qvel + wb * nv, nv, lane);
  warp_copy(cdof, d.cdof + wb * 6 * nv, 6 * nv, lane);
  __syncwarp();

  // ------------------------------------------------------------------ fwd_velocity
  const bool v_all = mask & STG_VELOCITY;  // the sub-stage bits serve the individually callable com_vel / passive / rne
  if (mask & (STG_VELOCITY | STG_COMVEL | STG_PASSIVE | STG_RNE)) {
    if (v_all || (mask & STG_RNE)) warp_copy(cinert, d.cinert + wb * 10 * nb, 10 * nb, lane);
    if (v_all)
#pragma unroll 1
    for (int a = lane; a < nu; a += 32) {  // actuator velocity = moment . qvel
      const int nnz = d.moment_rownnz[wb * nu + a], adr = d.moment_rowadr[wb * nu + a];
      float vel = 0.f;
      for (int k = 0; k < nnz; k++) vel += d.actuator_moment[wb * m.nJmom + adr + k] * qvel[d.moment_colind[wb * m.nJmom + adr + k]];
      d.actuator_velocity[wb * nu + a] = vel;
    }
    // com_vel: level-synchronous forward pass
    if (v_all || (mask & STG_COMVEL)) {
    if (lane < 6) cvel[lane] = 0.f;
    __syncwarp();
#pragma unroll 1
    for (int l = 1; l < m.nlevel; l++) {
#pragma unroll 1
      for (int i = m.level_adr[l] + lane; i < m.level_adr[l + 1]; i += 32) {
        const int b = m.level_body[i], pid = m.body_parentid[b], jntadr = m.body_jntadr[b], jntnum = m.body_jntnum[b];
        int dof = m.body_dofadr[b];
        float cv[6];
#pragma unroll
        for (int k = 0; k < 6; k++) cv[k] = cvel[6 * pid + k];
#pragma unroll 1
        for (int j = jntadr; j < jntadr + jntnum; j++) {
          const int t = m.jnt_type[j];
          if (t == JNT_FREE) {
            for (int q = 0; q < 3; q++) { const float v = qvel[dof + q]; for (int k = 0; k < 6; k++) { cv[k] += cdof[6 * (dof + q) + k] * v; cdofdot[6 * (dof + q) + k] = 0.f; } }
            for (int q = 3; q < 6; q++) motion_cross(cv, cdof + 6 * (dof + q), cdofdot + 6 * (dof + q));
            for (int q = 3; q < 6; q++) { const float v = qvel[dof + q]; for (int k = 0; k < 6; k++) cv[k] += cdof[6 * (dof + q) + k] * v; }
            dof += 6;
          } else if (t == JNT_BALL) {
            for (int q = 0; q < 3; q++) motion_cross(cv, cdof + 6 * (dof + q), cdofdot + 6 * (dof + q));
            for (int q = 0; q < 3; q++) { const float v = qvel[dof + q]; for (int k = 0; k < 6; k++) cv[k] += cdof[6 * (dof + q) + k] * v; }
            dof += 3;
          } else {
            motion_cross(cv, cdof + 6 * dof, cdofdot + 6 * dof);
            const float v = qvel[dof];
            for (int k = 0; k < 6; k++) cv[k] += cdof[6 * dof + k] * v;
            dof += 1;
          }
        }
#pragma unroll
        for (int k = 0; k < 6; k++) cvel[6 * b + k] = cv[k];
      }
      __syncwarp();
    }
    warp_copy(d.cvel + wb * 6 * nb, cvel, 6 * nb, lane);
    warp_copy(d.cdof_dot + wb * 6 * nv, cdofdot, 6 * nv, lane);
    } else if (mask & STG_RNE) {
      warp_copy(cvel, d.cvel + wb * 6 * nb, 6 * nb, lane);
      warp_copy(cdofdot, d.cdof_dot + wb * 6 * nv, 6 * nv, lane);
      __syncwarp();
    }

    // passive: joint springs (slide / hinge; ball and free joints through quat_sub) and dampers, gravity compensation
    if (v_all || (mask & STG_PASSIVE)) {
      const bool dsbl_spring = m.disableflags & DSBL_SPRING, dsbl_damper = m.disableflags & DSBL_DAMPER;
      const bool gravcomp = PEXT && m.has_gravcomp && !(m.disableflags & DSBL_GRAVITY) && !(dsbl_spring && dsbl_damper);
#pragma unroll 1
      for (int dd = lane; dd < nv; dd += 32) {
        const int j = m.dof_jntid[dd], t = m.jnt_type[j];
        float spring = 0.f, damper = 0.f, gc = 0.f;
        if (!(dsbl_spring && dsbl_damper)) {
          const float stiffness = m.jnt_stiffness[j];
          if (stiffness != 0.f && !dsbl_spring) {
            const int qa = m.jnt_qposadr[j], k = dd - m.jnt_dofadr[j];
            if (t == JNT_SLIDE || t == JNT_HINGE) spring = -(d.qpos[wb * m.nq + qa] - m.qpos_spring[qa]) * stiffness;
            else if (!PEXT) {}
            else if (t == JNT_FREE && k < 3) spring = -stiffness * (d.qpos[wb * m.nq + qa + k] - m.qpos_spring[qa + k]);
            else {  // rotational part: -k * quat_sub(q, q_spring) (passive.py:141-183, math.py:161-186)
              const int ra = t == JNT_FREE ? qa + 3 : qa, kk = t == JNT_FREE ? k - 3 : k;
              const q4 rot = qnormalize(ldq(d.qpos + wb * m.nq + ra)), ref = ldq(m.qpos_spring + ra);
              const q4 qd = qmul(mkq(ref.w, -ref.x, -ref.y, -ref.z), rot);
              const float s2 = sqrtf(qd.x * qd.x + qd.y * qd.y + qd.z * qd.z);
              if (s2 != 0.f) {
                float speed = 2.0f * atan2f(s2, qd.w);
                if (speed > 3.14159265358979f) speed -= 2.0f * 3.14159265358979f;
                spring = -stiffness * (kk == 0 ? qd.x : kk == 1 ? qd.y : qd.z) * (speed / s2);
              }
            }
          }
          const float damping = m.dof_damping[dd];
          if (damping != 0.f && !dsbl_damper) damper = -qvel[dd] * damping;
        }
        if (gravcomp) {  // passive.py:275-303: -gravity * mass * gravcomp at the body's inertial origin, projected on this dof
#pragma unroll 1
          for (int b = 1; b < nb; b++) {
            const float g = m.body_gravcomp[b];
            if (g == 0.f || !m.body_isdofancestor[b * nv + dd]) continue;
            const float sc = -m.body_mass[b] * g;
            const v3 off = ld3(d.xipos + (wb * nb + b) * 3) - ld3(d.subtree_com + (wb * nb + m.body_rootid[b]) * 3);
            const v3 jp = ld3(cdof + 6 * dd + 3) + cross(ld3(cdof + 6 * dd), off);
            gc += sc * (jp.x * m.gravity_x + jp.y * m.gravity_y + jp.z * m.gravity_z);
          }
        }
        d.qfrc_spring[wb * nv + dd] = spring;
        d.qfrc_damper[wb * nv + dd] = damper;
        d.qfrc_gravcomp[wb * nv + dd] = gc;
        const float passive = spring + damper + (m.jnt_actgravcomp[j] ? 0.f : gc);
        q_passive[dd] = passive;
        d.qfrc_passive[wb * nv + dd] = passive;
      }
    }
    // rne: cacc forward, cfrc per body, backward accumulation, projection
    if (v_all || (mask & STG_RNE)) {
    if (lane < 6) cacc[lane] = lane < 3 ? 0.f : ((m.disableflags & DSBL_GRAVITY) ? 0.f : -(lane == 3 ? m.gravity_x : lane == 4 ? m.gravity_y : m.gravity_z));
    __syncwarp();
#pragma unroll 1
    for (int l = 1; l < m.nlevel; l++) {
#pragma unroll 1
      for (int i = m.level_adr[l] + lane; i < m.level_adr[l + 1]; i += 32) {
        const int b = m.level_body[i], pid = m.body_parentid[b];
        float a[6];
#pragma unroll
        for (int k = 0; k < 6; k++) a[k] = cacc[6 * pid + k];
#pragma unroll 1
        for (int q = 0; q < m.body_dofnum[b]; q++) {
          const int dof = m.body_dofadr[b] + q;
          const float v = qvel[dof];
#pragma unroll
          for (int k = 0; k < 6; k++) a[k] += cdofdot[6 * dof + k] * v;
        }
#pragma unroll
        for (int k = 0; k < 6; k++) cacc[6 * b + k] = a[k];
      }
      __syncwarp();
    }
#pragma unroll 1
    for (int b = lane; b < nb; b += 32) {
      float f[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (b > 0) {
        float iv[6], g[6];
        inert_vec(cinert + 10 * b, cacc + 6 * b, f);
        inert_vec(cinert + 10 * b, cvel + 6 * b, iv);
        motion_cross_force(cvel + 6 * b, iv, g);
#pragma unroll
        for (int k = 0; k < 6; k++) f[k] += g[k];
      }
#pragma unroll
      for (int k = 0; k < 6; k++) cfrc[6 * b + k] = f[k];
    }
    __syncwarp();
#pragma unroll 1
    for (int l = m.nlevel - 2; l >= 0; l--) {
#pragma unroll 1
      for (int i = m.level_adr[l] + lane; i < m.level_adr[l + 1]; i += 32) {
        const int b = m.level_body[i];
        float acc[6];
#pragma unroll
        for (int k = 0; k < 6; k++) acc[k] = cfrc[6 * b + k];
#pragma unroll 1
        for (int c = m.body_childadr[b]; c < m.body_childadr[b + 1]; c++) {
          const float* cc = cfrc + 6 * m.body_childid[c];
#pragma unroll
          for (int k = 0; k < 6; k++) acc[k] += cc[k];
        }
#pragma unroll
        for (int k = 0; k < 6; k++) cfrc[6 * b + k] = acc[k];
      }
      __syncwarp();
    }
#pragma unroll 1
    for (int dd = lane; dd < nv; dd += 32) {
      const float v = dot6(cdof + 6 * dd, cfrc + 6 * m.dof_bodyid[dd]);
      q_bias[dd] = v;
      d.qfrc_bias[wb * nv + dd] = v;
    }
    warp_copy(d.cacc + wb * 6 * nb, cacc, 6 * nb, lane);
    warp_copy(d.cfrc_int + wb * 6 * nb, cfrc, 6 * nb, lane);
    }
  } else if (mask & STG_ACCELERATION) {
    warp_copy(q_passive, d.qfrc_passive + wb * nv, nv, lane);
    warp_copy(q_bias, d.qfrc_bias + wb * nv, nv, lane);
  }
  __syncwarp();

  // ------------------------------------------------------------------ fwd_actuation
  if (mask & STG_ACTUATION) {
#pragma unroll 1
    for (int dd = lane; dd < nv; dd += 32) q_act[dd] = 0.f;
    __syncwarp();
    const bool enabled = nu > 0 && !(m.disableflags & DSBL_ACTUATION);
    // actuator forces; scatter moment^T force by a per-dof gather loop over actuators (deterministic, no atomics)
#pragma unroll 1
    for (int a = lane; a < nu; a += 32) {
      float force = 0.f;
      if (enabled) {
        float ctrl = d.ctrl[wb * nu + a];
        if (m.actuator_ctrllimited[a] && !(m.disableflags & DSBL_CLAMPCTRL)) ctrl = clampf(ctrl, m.actuator_ctrlrange[2 * a], m.actuator_ctrlrange[2 * a + 1]);
        const float length = d.actuator_length[wb * nu + a], velocity = d.actuator_velocity[wb * nu + a];
        const float *gp = m.actuator_gainprm + 10 * a, *bp = m.actuator_biasprm + 10 * a;
        float gain = 0.f, bias = 0.f;
        if (m.actuator_gaintype[a] == GAIN_FIXED) gain = gp[0];
        else if (m.actuator_gaintype[a] == GAIN_AFFINE) gain = gp[0] + gp[1] * length + gp[2] * velocity;
        if (m.actuator_biastype[a] == BIAS_AFFINE) bias = bp[0] + bp[1] * length + bp[2] * velocity;
        force = gain * ctrl + bias;
        if (m.actuator_forcelimited[a]) force = clampf(force, m.actuator_forcerange[2 * a], m.actuator_forcerange[2 * a + 1]);
      }
      d.actuator_force[wb * nu + a] = force;
      aforce[a] = force;
    }
    __syncwarp();
    if (enabled) {
#pragma unroll 1
      for (int dd = lane; dd < nv; dd += 32) {
        float q = 0.f;
        // moment^T force through the per-dof reverse table (entries in actuator order -> fixed summation order, no atomics)
        for (int k = m.dofact_adr[dd]; k < m.dofact_adr[dd + 1]; k++) q += d.actuator_moment[wb * m.nJmom + m.dofact_mom[k]] * aforce[m.dofact_act[k]];
        const int j = m.dof_jntid[dd];
        if (!(m.disableflags & DSBL_GRAVITY) && m.jnt_actgravcomp[j]) q += d.qfrc_gravcomp[wb * nv + dd];
        if (m.jnt_actfrclimited[j]) q = clampf(q, m.jnt_actfrcrange[2 * j], m.jnt_actfrcrange[2 * j + 1]);
        q_act[dd] = q;
      }
    }
    __syncwarp();
    warp_copy(d.qfrc_actuator + wb * nv, q_act, nv, lane);
  } else if (mask & STG_ACCELERATION) {
    warp_copy(q_act, d.qfrc_actuator + wb * nv, nv, lane);
  }
  __syncwarp();

  // ------------------------------------------------------------------ fwd_acceleration (factorize=True)
  if (mask & (STG_ACCELERATION | STG_FACTOR_ONLY)) {
    if (mask & STG_ACCELERATION) {
#pragma unroll 1
      for (int dd = lane; dd < nv; dd += 32) q_smooth[dd] = q_passive[dd] - q_bias[dd] + q_act[dd] + d.qfrc_applied[wb * nv + dd];
      // xfrc_applied: skipped entirely when the world's applied wrenches are all zero (the common case)
      bool any = false;
#pragma unroll 1
      for (int i = lane; i < 6 * nb; i += 32) any |= d.xfrc_applied[wb * 6 * nb + i] != 0.f;
      if (__any_sync(FULL_MASK, any)) {
#pragma unroll 1
        for (int dd = lane; dd < nv; dd += 32) {
          const float* cd = cdof + 6 * dd;
          const int db = m.dof_bodyid[dd];
          float acc = 0.f;
          for (int b = db; b < nb; b++) {
            const float* ft = d.xfrc_applied + (wb * nb + b) * 6;
            if (ft[0] == 0.f && ft[1] == 0.f && ft[2] == 0.f && ft[3] == 0.f && ft[4] == 0.f && ft[5] == 0.f) continue;
            int p = b;
            while (p != 0 && p != db) p = m.body_parentid[p];
            if (p == 0) continue;
            const v3 off = ld3(d.xipos + (wb * nb + b) * 3) - ld3(d.subtree_com + (wb * nb + m.body_rootid[b]) * 3);
            const v3 cr = cross(ld3(cd), off);
            acc += cd[3] * ft[0] + cd[4] * ft[1] + cd[5] * ft[2] + cd[0] * ft[3] + cd[1] * ft[4] + cd[2] * ft[5] + dot(cr, ld3(ft));
          }
          q_smooth[dd] += acc;
        }
      }
      __syncwarp();
      warp_copy(d.qfrc_smooth + wb * nv, q_smooth, nv, lane);
    }
    // per-tree dense Cholesky of M, qLD block = upper factor U (row-major, zeros below), qacc_smooth = M^-1 qfrc_smooth
    const float* Mw = d.M + wb * m.nC;
#pragma unroll 1
    for (int t = 0; t < m.ntree; t++) {
      const int start = m.tree_dofadr[t], n = m.tree_dofnum[t], ld = chol_ld(n);
#pragma unroll 1
      for (int i = lane; i < n * ld; i += 32) A[i] = 0.f;
      __syncwarp();
      const int e0 = m.M_rowadr[start], e1 = m.M_rowadr[start + n - 1] + m.M_rownnz[start + n - 1];
#pragma unroll 1
      for (int e = e0 + lane; e < e1; e += 32) A[(m.M_entry_row[e] - start) * ld + (m.M_colind[e] - start)] = Mw[e];
      __syncwarp();
      float* qld = d.qLD + wb * m.qld_total + m.tree_qLDadr[t];
      if (n <= 32) {
        const float b = ((mask & STG_ACCELERATION) && lane < n) ? q_smooth[start + lane] : 0.f;
        const float xx = chol_solve_reg_any(A, ld, n, b, A, ld, lane);
        __syncwarp();
        if ((mask & STG_ACCELERATION) && lane < n) d.qacc_smooth[wb * nv + start + lane] = xx;
      } else {
        warp_cholesky(A, n, ld, lane);
        if (mask & STG_ACCELERATION) {
#pragma unroll 1
          for (int i = lane; i < n; i += 32) x[i] = q_smooth[start + i];
          __syncwarp();
          warp_chol_solve(A, n, ld, x, lane);
#pragma unroll 1
          for (int i = lane; i < n; i += 32) d.qacc_smooth[wb * nv + start + i] = x[i];
        }
      }
#pragma unroll 1
      for (int r = 0; r < n; r++)
        for (int c = lane; c < n; c += 32) qld[r * n + c] = c >= r ? A[c * ld + r] : 0.f;
      __syncwarp();
    }
  }
}

}  // namespace

size_t smem_velocity(const ModelDev& m) { return (size_t)vel_layout(m).total * sizeof(float) * MJB_WARPS_PER_BLOCK; }

cudaError_t launch_velocity(const ModelDev& m, const DataDev& d, int mask, cudaStream_t s) {
  const size_t smem = smem_velocity(m);
  static size_t configured[2] = {0, 0};
  const int ext = m.has_gravcomp ? 1 : 0;  // has_gravcomp also flags free / ball joint springs (io.py put_model)
  void (*kern)(ModelDev, DataDev, int) = ext ? k_velocity<true> : k_velocity<false>;
  if (smem > 48 * 1024 && smem > configured[ext]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[ext] = smem;
  }
  const int grid = d.wn;
  kern<<<grid, MJB_WARPS_PER_BLOCK * 32, smem, s>>>(m, d, mask);
  return cudaGetLastError();
}
