// dex_kin.cuh — warp-per-env forward kinematics (restates mj_kinematics).
// Lanes map to bodies; the tree is walked level by level (depth <= model.maxdepth) with a
// warp barrier between levels.  All outputs live in the calling warp's shared-memory slice.
#ifndef DEX_KIN_CUH_
#define DEX_KIN_CUH_

#include "dex_model.cuh"

template <typename T>
struct KinBuf {
  T* xpos;     // [MAXBODY*3]
  T* xquat;    // [MAXBODY*4]
  T* xmat;     // [MAXBODY*9]
  T* xanchor;  // [MAXJNT*3] (may be null)
  T* xaxis;    // [MAXJNT*3] (may be null)
};

template <typename T>
DEX_DEV void dex_fk(const DevModel<T>& m, const T* qpos, const KinBuf<T>& k) {
  const int lane = dex_lane();
  for (int L = 0; L <= m.maxdepth; ++L) {
    for (int b = lane; b < m.nbody; b += 32) {
      if (dex_ldg(m.body_depth + b) != L) continue;
      T pos[3], quat[4], mat[9], t[3];
      if (b == 0) {
        pos[0] = pos[1] = pos[2] = 0;
        quat[0] = 1; quat[1] = quat[2] = quat[3] = 0;
      } else {
        const int p = dex_ldg(m.body_parent + b);
        T bp[3] = {dex_ldg(m.body_pos + 3 * b), dex_ldg(m.body_pos + 3 * b + 1), dex_ldg(m.body_pos + 3 * b + 2)};
        T bq[4] = {dex_ldg(m.body_quat + 4 * b), dex_ldg(m.body_quat + 4 * b + 1), dex_ldg(m.body_quat + 4 * b + 2),
                   dex_ldg(m.body_quat + 4 * b + 3)};
        rot_vec(t, k.xmat + 9 * p, bp);
        pos[0] = k.xpos[3 * p] + t[0]; pos[1] = k.xpos[3 * p + 1] + t[1]; pos[2] = k.xpos[3 * p + 2] + t[2];
        quat_mul(quat, k.xquat + 4 * p, bq);
        const int ja = dex_ldg(m.body_jntadr + b), jn = dex_ldg(m.body_jntnum + b);
        for (int j = ja; j < ja + jn; ++j) {
          const int qa = dex_ldg(m.jnt_qposadr + j);
          T ax[3] = {dex_ldg(m.jnt_axis + 3 * j), dex_ldg(m.jnt_axis + 3 * j + 1), dex_ldg(m.jnt_axis + 3 * j + 2)};
          if (dex_ldg(m.jnt_type + j) == DEX_JNT_FREE) {
            pos[0] = qpos[qa]; pos[1] = qpos[qa + 1]; pos[2] = qpos[qa + 2];
            quat[0] = qpos[qa + 3]; quat[1] = qpos[qa + 4]; quat[2] = qpos[qa + 5]; quat[3] = qpos[qa + 6];
            quat_normalize(quat);
            if (k.xanchor) {
              quat2mat(mat, quat);
              k.xanchor[3 * j] = pos[0]; k.xanchor[3 * j + 1] = pos[1]; k.xanchor[3 * j + 2] = pos[2];
              rot_vec(k.xaxis + 3 * j, mat, ax);
            }
          } else {
            T jp[3] = {dex_ldg(m.jnt_pos + 3 * j), dex_ldg(m.jnt_pos + 3 * j + 1), dex_ldg(m.jnt_pos + 3 * j + 2)};
            T anchor[3];
            quat2mat(mat, quat);
            rot_vec(t, mat, jp);
            anchor[0] = pos[0] + t[0]; anchor[1] = pos[1] + t[1]; anchor[2] = pos[2] + t[2];
            if (k.xanchor) {
              k.xanchor[3 * j] = anchor[0]; k.xanchor[3 * j + 1] = anchor[1]; k.xanchor[3 * j + 2] = anchor[2];
              rot_vec(k.xaxis + 3 * j, mat, ax);
            }
            T ql[4], qn[4];
            axisangle2quat(ql, ax, qpos[qa] - dex_ldg(m.qpos0 + qa));
            quat_mul(qn, quat, ql);
            quat[0] = qn[0]; quat[1] = qn[1]; quat[2] = qn[2]; quat[3] = qn[3];
            quat2mat(mat, quat);
            rot_vec(t, mat, jp);
            pos[0] = anchor[0] - t[0]; pos[1] = anchor[1] - t[1]; pos[2] = anchor[2] - t[2];
          }
        }
      }
      quat_normalize(quat);
      quat2mat(mat, quat);
      for (int i = 0; i < 3; ++i) k.xpos[3 * b + i] = pos[i];
      for (int i = 0; i < 4; ++i) k.xquat[4 * b + i] = quat[i];
      for (int i = 0; i < 9; ++i) k.xmat[9 * b + i] = mat[i];
    }
    dex_syncwarp();
  }
}

// cdof of dof `lane` (mj_comPos): [angular(3); linear(3)] about the reference point of its tree.
// hinge: [axis; axis x (ref - anchor)];  free: 3 world translations then 3 body-frame rotations.
template <typename T>
DEX_DEV void dex_cdof_lane(const DevModel<T>& m, int lane, const T* xpos, const T* xmat, const T* xanchor,
                           const T* xaxis, const T* ref, T* cd) {
  for (int k = 0; k < 6; ++k) cd[k] = 0;
  if (lane >= m.nv) return;
  const int j = dex_ldg(m.dof_jnt + lane);
  const int b = dex_ldg(m.dof_body + lane);
  const T* rp = ref + 3 * dex_ldg(m.dof_tree + lane);
  if (dex_ldg(m.jnt_type + j) == DEX_JNT_FREE) {
    const int k = lane - dex_ldg(m.jnt_dofadr + j);
    if (k < 3) cd[3 + k] = 1;
    else {
      const T* R = xmat + 9 * b;
      T ax[3] = {R[k - 3], R[3 + k - 3], R[6 + k - 3]};
      T off[3] = {rp[0] - xpos[3 * b], rp[1] - xpos[3 * b + 1], rp[2] - xpos[3 * b + 2]};
      cd[0] = ax[0]; cd[1] = ax[1]; cd[2] = ax[2];
      cross3(cd + 3, ax, off);
    }
  } else {
    T ax[3] = {xaxis[3 * j], xaxis[3 * j + 1], xaxis[3 * j + 2]};
    T off[3] = {rp[0] - xanchor[3 * j], rp[1] - xanchor[3 * j + 1], rp[2] - xanchor[3 * j + 2]};
    cd[0] = ax[0]; cd[1] = ax[1]; cd[2] = ax[2];
    cross3(cd + 3, ax, off);
  }
}

#endif  // DEX_KIN_CUH_
