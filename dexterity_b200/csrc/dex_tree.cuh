// dex_tree.cuh — K0: THREAD-per-env tree kernel (restates SURVEY.md §8a rows 10-13 and 16-18):
//   mj_kinematics, mj_comPos (cinert / cdof about a per-tree reference point), mj_comVel, mj_rne
//   (bias; constant xfrc_applied folded in as external body forces), mj_crb (mass matrix), mj_passive
//   (joint damping), mj_fwdActuation -> qfrc_smooth.
//
// Why one thread per env here (and a warp per env for collision / the solver): the kinematic tree is
// the same in every env, so 32 envs walk it in lock step with NO divergence, while a warp that maps
// lanes to bodies keeps 2-6 of 32 lanes busy per tree level (r01f profile: FK twice + the tree passes
// cost 13 k warp instructions per env, 22 % of the step; this kernel needs ~0.5 k per env).
//
// The tree is walked once, depth first (bodies are stored in DFS pre-order, verified by the host
// model builder), with a stack indexed by depth: entering a body computes its pose, cdof, cinert,
// cvel / cacc and its own spatial force; leaving it (all descendants done) yields the bias force and
// the mass-matrix rows of its dofs from the accumulated composite inertia / force, which are then
// folded into the parent's stack slot.  Per-thread state is ~3 KB of local memory (L1 resident).
//
// Outputs (global scratch, consumed by K1 / K2):
//   xkin [env][body][16]  : xpos(3) xquat(4) xmat(9)
//   tree [env][8][32]     : rows 0-5 cdof component k of dof (column), row 6 qfrc_smooth, row 7 tree reference points
//   mscr [env][32][32]    : dense symmetric M (only the structural non-zeros are ever written; the
//                           buffer is zero-initialised once by the caller)
#ifndef DEX_TREE_CUH_
#define DEX_TREE_CUH_

#include "dex_model.cuh"

#ifndef DEX_MAXDEPTH
#define DEX_MAXDEPTH 12
#endif
#define DEX_BODY_REC 24   // packed per-body constants: pos(3) mass | quat(4) | ipos(3) - | inertia6 | xfrc6
#define DEX_XKIN_BODY 16
#define DEX_XKIN_ENV (DEX_MAXBODY * DEX_XKIN_BODY)
#define DEX_TREE_ENV 256
#define DEX_TREE_ROW_FS 6
#define DEX_TREE_ROW_REF 7
#define DEX_MG_ENV (32 * 32)                  // global M scratch: [env][32][32] (symmetric, row j = M[j][:])

// Pose of body b from its parent's pose (mj_kinematics for one body).  When `ref` is given, the cdof
// columns of the body's dofs are written to cdof[dof][6] ([angular; linear] about ref[tree]).
template <typename T>
DEX_DEV void dex_body_fk(const DevModel<T>& m, const T* qpos, int b, int ja, int jn, int tree, const T* ppos, const T* pquat,
                         const T* pmat, T* pos, T* quat, T* mat, const T* ref, T (*cdof)[6]) {
  T t[3];
  T bp[4], bq[4];   // packed body record (DevModel::body_rec): pos(3) mass | quat(4) | ...
  dex_ldg4(m.body_rec + DEX_BODY_REC * b, bp);
  dex_ldg4(m.body_rec + DEX_BODY_REC * b + 4, bq);
  rot_vec(t, pmat, bp);
  pos[0] = ppos[0] + t[0]; pos[1] = ppos[1] + t[1]; pos[2] = ppos[2] + t[2];
  quat_mul(quat, pquat, bq);
  const T* rp = ref ? ref + 3 * tree : nullptr;
  int freedof = -1;
  for (int j = ja; j < ja + jn; ++j) {
    int ji[4];   // packed joint records (DevModel::jnt_irec / jnt_rec): type, qposadr, dofadr | axis(3) qpos0 | pos(3)
    T ax[4], jp[4];
    dex_ldg_int4(m.jnt_irec + 4 * j, ji);
    dex_ldg4(m.jnt_rec + 8 * j, ax);
    dex_ldg4(m.jnt_rec + 8 * j + 4, jp);
    const int qa = ji[1];
    if (ji[0] == DEX_JNT_FREE) {
      pos[0] = qpos[qa]; pos[1] = qpos[qa + 1]; pos[2] = qpos[qa + 2];
      quat[0] = qpos[qa + 3]; quat[1] = qpos[qa + 4]; quat[2] = qpos[qa + 5]; quat[3] = qpos[qa + 6];
      quat_normalize(quat);
      freedof = ji[2];
    } else {
      T anchor[3];
      quat2mat(mat, quat);
      rot_vec(t, mat, jp);
      anchor[0] = pos[0] + t[0]; anchor[1] = pos[1] + t[1]; anchor[2] = pos[2] + t[2];
      if (rp) {  // hinge: [axis; axis x (ref - anchor)]
        T* cd = cdof[ji[2]];
        T xa[3];
        rot_vec(xa, mat, ax);
        const T off[3] = {rp[0] - anchor[0], rp[1] - anchor[1], rp[2] - anchor[2]};
        cd[0] = xa[0]; cd[1] = xa[1]; cd[2] = xa[2];
        cross3(cd + 3, xa, off);
      }
      T ql[4], qn[4];
      axisangle2quat(ql, ax, qpos[qa] - ax[3]);
      quat_mul(qn, quat, ql);
      quat[0] = qn[0]; quat[1] = qn[1]; quat[2] = qn[2]; quat[3] = qn[3];
      quat2mat(mat, quat);
      rot_vec(t, mat, jp);
      pos[0] = anchor[0] - t[0]; pos[1] = anchor[1] - t[1]; pos[2] = anchor[2] - t[2];
    }
  }
  quat_normalize(quat);
  quat2mat(mat, quat);
  if (rp && freedof >= 0) {  // free joint: 3 world translations, then 3 body-frame rotations
    const T off[3] = {rp[0] - pos[0], rp[1] - pos[1], rp[2] - pos[2]};
    for (int k = 0; k < 3; ++k) {
      T* ct = cdof[freedof + k];
      for (int i = 0; i < 6; ++i) ct[i] = 0;
      ct[3 + k] = 1;
      T* cr = cdof[freedof + 3 + k];
      const T axk[3] = {mat[k], mat[3 + k], mat[6 + k]};
      cr[0] = axk[0]; cr[1] = axk[1]; cr[2] = axk[2];
      cross3(cr + 3, axk, off);
    }
  }
}

template <typename T>
DEX_DEV void dex_tree_env(const DevModel<T>& m, const BatchState<T>& st, int env) {
  if (dex_env_off(st, env)) return;
  const int nv = m.nv, nb = m.nbody;
  T qpos[48], qvel[DEX_MAXDOF];
  for (int i = 0; i < m.nq; ++i) qpos[i] = st.qpos[(size_t)env * st.qpos_stride + i];
  for (int i = 0; i < nv; ++i) qvel[i] = st.qvel[(size_t)env * st.qvel_stride + i];
  T* xkin = st.xkin + (size_t)env * DEX_XKIN_ENV;
  T* tree = st.tree + (size_t)env * DEX_TREE_ENV;
  T* M = st.mscr + (size_t)env * DEX_MG_ENV;

  // ---- per-tree reference points: FK along the path world -> reference body
  T ref[16];
  for (int k = 0; k < 16; ++k) ref[k] = 0;
  for (int t = 1; t < m.ntree; ++t) {
    int path[DEX_MAXDEPTH], n = 0;
    for (int b = dex_ldg(m.tree_refbody + t); b > 0; b = dex_ldg(m.body_parent + b)) path[n++] = b;
    T pos[3] = {0, 0, 0}, quat[4] = {1, 0, 0, 0}, mat[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = n - 1; i >= 0; --i) {
      T npos[3], nquat[4], nmat[9];
      dex_body_fk<T>(m, qpos, path[i], dex_ldg(m.body_jntadr + path[i]), dex_ldg(m.body_jntnum + path[i]), 0, pos, quat, mat, npos,
                     nquat, nmat, nullptr, nullptr);
      for (int k = 0; k < 3; ++k) pos[k] = npos[k];
      for (int k = 0; k < 4; ++k) quat[k] = nquat[k];
      for (int k = 0; k < 9; ++k) mat[k] = nmat[k];
    }
    ref[3 * t] = pos[0]; ref[3 * t + 1] = pos[1]; ref[3 * t + 2] = pos[2];
  }

  // ---- actuation (mj_fwdActuation) and passive damping (before the walk: qfrc_smooth is finished as each body is left)
  T act[DEX_MAXDOF];
  for (int i = 0; i < DEX_MAXDOF; ++i) act[i] = 0;
  for (int a = 0; a < m.nu; ++a) {
    const int tid = dex_ldg(m.act_trnid + a);
    const T gear = dex_ldg(m.act_gear + a);
    const bool joint = dex_ldg(m.act_trntype + a) == 0;
    T length = 0, vel = 0;
    if (joint) {
      length = qpos[dex_ldg(m.jnt_qposadr + tid)];
      vel = qvel[dex_ldg(m.jnt_dofadr + tid)];
    } else {
      const int wa = dex_ldg(m.ten_adr + tid), wn = dex_ldg(m.ten_num + tid);
      for (int w = wa; w < wa + wn; ++w) {
        const int j = dex_ldg(m.wrap_jnt + w);
        const T c = dex_ldg(m.wrap_coef + w);
        length += c * qpos[dex_ldg(m.jnt_qposadr + j)];
        vel += c * qvel[dex_ldg(m.jnt_dofadr + j)];
      }
    }
    length *= gear; vel *= gear;
    T c = st.ctrl[(size_t)env * st.ctrl_stride + a];
    if (dex_ldg(m.act_ctrllimited + a)) c = dx_clamp(c, dex_ldg(m.act_ctrlrange + 2 * a), dex_ldg(m.act_ctrlrange + 2 * a + 1));
    T f = dex_ldg(m.act_gainprm + 3 * a) * c + dex_ldg(m.act_biasprm + 3 * a) + dex_ldg(m.act_biasprm + 3 * a + 1) * length +
          dex_ldg(m.act_biasprm + 3 * a + 2) * vel;
    if (dex_ldg(m.act_forcelimited + a)) f = dx_clamp(f, dex_ldg(m.act_forcerange + 2 * a), dex_ldg(m.act_forcerange + 2 * a + 1));
    if (joint) act[dex_ldg(m.jnt_dofadr + tid)] += gear * f;
    else {
      const int wa = dex_ldg(m.ten_adr + tid), wn = dex_ldg(m.ten_num + tid);
      for (int w = wa; w < wa + wn; ++w) act[dex_ldg(m.jnt_dofadr + dex_ldg(m.wrap_jnt + w))] += gear * dex_ldg(m.wrap_coef + w) * f;
    }
  }
  // ---- depth-first walk
  T s_pos[DEX_MAXDEPTH][3], s_quat[DEX_MAXDEPTH][4], s_mat[DEX_MAXDEPTH][9];
  T s_vel[DEX_MAXDEPTH][6], s_acc[DEX_MAXDEPTH][6], s_crb[DEX_MAXDEPTH][10], s_frc[DEX_MAXDEPTH][6];
  int s_body[DEX_MAXDEPTH];
  T cdof[DEX_MAXDOF][6];
  for (int k = 0; k < 3; ++k) s_pos[0][k] = 0;
  s_quat[0][0] = 1; s_quat[0][1] = s_quat[0][2] = s_quat[0][3] = 0;
  for (int k = 0; k < 9; ++k) s_mat[0][k] = (k % 4 == 0) ? T(1) : T(0);
  for (int k = 0; k < 6; ++k) { s_vel[0][k] = 0; s_acc[0][k] = 0; }
  s_acc[0][3] = -m.gravity[0]; s_acc[0][4] = -m.gravity[1]; s_acc[0][5] = -m.gravity[2];
  s_body[0] = 0;
  // world body record
  for (int k = 0; k < 3; ++k) xkin[k] = 0;
  for (int k = 0; k < 4; ++k) xkin[3 + k] = s_quat[0][k];
  for (int k = 0; k < 9; ++k) xkin[7 + k] = s_mat[0][k];

  int cur = 0;
  for (int b = 1; b <= nb; ++b) {
    const int d = b < nb ? dex_ldg(m.body_depth + b) : 1;
    // ---- leave every body at depth >= d: its subtree is complete
    for (; cur >= d; --cur) {
      const int e = s_body[cur];
      const int da = dex_ldg(m.body_dofadr + e), dn = dex_ldg(m.body_dofnum + e);
      for (int i = da; i < da + dn; ++i) {
        const T* ci = cdof[i];
        const T* cf = s_frc[cur];
        const T bias = ci[0] * cf[0] + ci[1] * cf[1] + ci[2] * cf[2] + ci[3] * cf[3] + ci[4] * cf[4] + ci[5] * cf[5];
        tree[DEX_TREE_ROW_FS * 32 + i] = -dex_ldg(m.dof_damping + i) * qvel[i] - bias + act[i];  // qfrc_smooth
        for (int k = 0; k < 6; ++k) tree[k * 32 + i] = ci[k];
        T buf[6];
        mul_inert_vec(buf, s_crb[cur], ci);
        for (int j = i; j >= 0; j = dex_ldg(m.dof_parent + j)) {
          const T* cj = cdof[j];
          T s = cj[0] * buf[0] + cj[1] * buf[1] + cj[2] * buf[2] + cj[3] * buf[3] + cj[4] * buf[4] + cj[5] * buf[5];
          if (j == i) s += dex_ldg(m.dof_armature + i);
          M[i * 32 + j] = s;
          M[j * 32 + i] = s;
        }
      }
      if (cur > 1) {
        for (int k = 0; k < 6; ++k) s_frc[cur - 1][k] += s_frc[cur][k];
        for (int k = 0; k < 10; ++k) s_crb[cur - 1][k] += s_crb[cur][k];
      }
    }
    if (b == nb) break;
    // ---- enter body b at depth d
    cur = d;
    s_body[d] = b;
    T* pos = s_pos[d];
    T* mat = s_mat[d];
    dex_body_fk<T>(m, qpos, b, dex_ldg(m.body_jntadr + b), dex_ldg(m.body_jntnum + b), dex_ldg(m.body_tree + b), s_pos[d - 1],
                   s_quat[d - 1], s_mat[d - 1], pos, s_quat[d], mat, ref, cdof);
    {
      const T rec[16] = {pos[0], pos[1], pos[2], s_quat[d][0], s_quat[d][1], s_quat[d][2], s_quat[d][3], mat[0], mat[1], mat[2],
                         mat[3], mat[4], mat[5], mat[6], mat[7], mat[8]};
      dex_store16(xkin + b * DEX_XKIN_BODY, rec);
    }
    // cinert about the tree reference point (mj_comPos)
    const T* rp = ref + 3 * dex_ldg(m.body_tree + b);
    const T mass = dex_ldg(m.body_mass + b);
    const T ip[3] = {dex_ldg(m.body_ipos + 3 * b), dex_ldg(m.body_ipos + 3 * b + 1), dex_ldg(m.body_ipos + 3 * b + 2)};
    T off[3];
    rot_vec(off, mat, ip);
    off[0] += pos[0] - rp[0]; off[1] += pos[1] - rp[1]; off[2] += pos[2] - rp[2];
    {
      const T* I6 = m.body_inertia6 + 6 * b;
      const T Ib[9] = {dex_ldg(I6), dex_ldg(I6 + 3), dex_ldg(I6 + 4), dex_ldg(I6 + 3), dex_ldg(I6 + 1), dex_ldg(I6 + 5),
                       dex_ldg(I6 + 4), dex_ldg(I6 + 5), dex_ldg(I6 + 2)};
      T RI[9];
      for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc)
        RI[3 * r + cc] = mat[3 * r] * Ib[cc] + mat[3 * r + 1] * Ib[3 + cc] + mat[3 * r + 2] * Ib[6 + cc];
      T* c = s_crb[d];
      const T o2 = dot3(off, off);
      c[0] = RI[0] * mat[0] + RI[1] * mat[1] + RI[2] * mat[2] + mass * (o2 - off[0] * off[0]);
      c[1] = RI[3] * mat[3] + RI[4] * mat[4] + RI[5] * mat[5] + mass * (o2 - off[1] * off[1]);
      c[2] = RI[6] * mat[6] + RI[7] * mat[7] + RI[8] * mat[8] + mass * (o2 - off[2] * off[2]);
      c[3] = RI[0] * mat[3] + RI[1] * mat[4] + RI[2] * mat[5] - mass * off[0] * off[1];
      c[4] = RI[0] * mat[6] + RI[1] * mat[7] + RI[2] * mat[8] - mass * off[0] * off[2];
      c[5] = RI[3] * mat[6] + RI[4] * mat[7] + RI[5] * mat[8] - mass * off[1] * off[2];
      c[6] = mass * off[0]; c[7] = mass * off[1]; c[8] = mass * off[2]; c[9] = mass;
    }
    // velocity / acceleration through the body's joints (mj_comVel, mj_rne forward)
    T v[6], a[6];
    for (int k = 0; k < 6; ++k) { v[k] = s_vel[d - 1][k]; a[k] = s_acc[d - 1][k]; }
    {
      const int ja = dex_ldg(m.body_jntadr + b), jn = dex_ldg(m.body_jntnum + b);
      for (int j = ja; j < ja + jn; ++j) {
        const int da = dex_ldg(m.jnt_dofadr + j);
        if (dex_ldg(m.jnt_type + j) == DEX_JNT_FREE) {
          for (int i = 0; i < 3; ++i) for (int k = 0; k < 6; ++k) v[k] += cdof[da + i][k] * qvel[da + i];
          T cdd[3][6];
          for (int i = 0; i < 3; ++i) cross_motion(cdd[i], v, cdof[da + 3 + i]);
          for (int i = 0; i < 3; ++i) for (int k = 0; k < 6; ++k) {
            v[k] += cdof[da + 3 + i][k] * qvel[da + 3 + i];
            a[k] += cdd[i][k] * qvel[da + 3 + i];
          }
        } else {
          T cdd[6];
          cross_motion(cdd, v, cdof[da]);
          for (int k = 0; k < 6; ++k) { v[k] += cdof[da][k] * qvel[da]; a[k] += cdd[k] * qvel[da]; }
        }
      }
    }
    for (int k = 0; k < 6; ++k) { s_vel[d][k] = v[k]; s_acc[d][k] = a[k]; }
    // own spatial force: I a + v x* (I v) - external force about the reference point
    {
      const T* xf = m.body_xfrc + 6 * b;
      const T f[3] = {dex_ldg(xf), dex_ldg(xf + 1), dex_ldg(xf + 2)};
      T tq[3], t1[6], t2[6], t3[6];
      cross3(tq, off, f);
      mul_inert_vec(t1, s_crb[d], a);
      mul_inert_vec(t2, s_crb[d], v);
      cross_force(t3, v, t2);
      T* cf = s_frc[d];
      cf[0] = -(dex_ldg(xf + 3) + tq[0]) + t1[0] + t3[0];
      cf[1] = -(dex_ldg(xf + 4) + tq[1]) + t1[1] + t3[1];
      cf[2] = -(dex_ldg(xf + 5) + tq[2]) + t1[2] + t3[2];
      cf[3] = -f[0] + t1[3] + t3[3]; cf[4] = -f[1] + t1[4] + t3[4]; cf[5] = -f[2] + t1[5] + t3[5];
    }
  }

  for (int k = 0; k < 16; ++k) tree[DEX_TREE_ROW_REF * 32 + k] = ref[k];
}

// ------------------------------------------------------------------------------------------
// K0, second form: EIGHT LANES per env (four envs per warp), one lane per kinematic chain.
// The thread-per-env walk above is the most instruction-efficient form, but at 16 384 envs it leaves a
// B200 with 3.5 warps per SM and 3 KB of local state per thread, i.e. bound by L2 latency (r01g profile:
// IPC 0.25, 0.27 ms per launch).  Here the host assigns every body to a (round, lane) slot, one round per
// tree depth; a lane keeps the pose / velocity / acceleration of its current body in registers, a child
// fetches its parent's state with shuffles from the parent's lane, and the own inertia / force of every
// body lives in shared memory where children add their subtree totals on the way back up.  Same
// arithmetic as dex_tree_env, body by body.  Models whose widest tree level exceeds 8 bodies keep the
// thread-per-env kernel (DevModel::k0_rounds == 0).
// ------------------------------------------------------------------------------------------
#define DEX_K0_G 8
#define DEX_K0_ENVS 4
#define DEX_K0_OFF_OWN 0                                   // [body][16]: crb(10) cfrc(6)
#define DEX_K0_OFF_CDOF (DEX_K0_OFF_OWN + DEX_MAXBODY * 16)  // [dof][6]
#define DEX_K0_OFF_ACT (DEX_K0_OFF_CDOF + DEX_MAXDOF * 6)
#define DEX_K0_OFF_REF (DEX_K0_OFF_ACT + DEX_MAXDOF)
#define DEX_K0_ENV_T (DEX_K0_OFF_REF + 16)
#define DEX_K0_SMEM_T (DEX_K0_ENVS * DEX_K0_ENV_T)

template <typename T>
DEX_DEV void dex_tree_warp(const DevModel<T>& m, const BatchState<T>& st, int env_base, T* smem) {
  const int lane = dex_lane(), gl = lane & (DEX_K0_G - 1), ge = lane / DEX_K0_G;
  const int env = env_base + ge;
  const bool live = env < st.nenv && !dex_env_off(st, env < st.nenv ? env : st.nenv - 1);   // dead groups compute on a valid env and write nothing
  if (!dex_ballot(live)) return;   // nothing to do for any of the warp's envs (common in the placement rounds)
  const int envr = env < st.nenv ? env : st.nenv - 1;
  T* sm = smem + ge * DEX_K0_ENV_T;
  T* own = sm + DEX_K0_OFF_OWN;
  T* cdofs = sm + DEX_K0_OFF_CDOF;
  T (*cdof)[6] = reinterpret_cast<T (*)[6]>(cdofs);
  T* act = sm + DEX_K0_OFF_ACT;
  T* ref = sm + DEX_K0_OFF_REF;
  const T* qpos = st.qpos + (size_t)envr * st.qpos_stride;
  const T* qvel = st.qvel + (size_t)envr * st.qvel_stride;
  T* xkin = st.xkin + (size_t)envr * DEX_XKIN_ENV;
  T* tree = st.tree + (size_t)envr * DEX_TREE_ENV;
  T* M = st.mscr + (size_t)envr * DEX_MG_ENV;

  for (int i = gl; i < DEX_MAXDOF; i += DEX_K0_G) act[i] = 0;
  for (int i = gl; i < 16; i += DEX_K0_G) ref[i] = 0;
  dex_syncwarp();
  // ---- actuation (lane = actuator)
  for (int a = gl; a < m.nu; a += DEX_K0_G) {
    const int tid = dex_ldg(m.act_trnid + a);
    const T gear = dex_ldg(m.act_gear + a);
    const bool joint = dex_ldg(m.act_trntype + a) == 0;
    T length = 0, vel = 0;
    if (joint) {
      length = qpos[dex_ldg(m.jnt_qposadr + tid)];
      vel = qvel[dex_ldg(m.jnt_dofadr + tid)];
    } else {
      const int wa = dex_ldg(m.ten_adr + tid), wn = dex_ldg(m.ten_num + tid);
      for (int w = wa; w < wa + wn; ++w) {
        const int j = dex_ldg(m.wrap_jnt + w);
        const T c = dex_ldg(m.wrap_coef + w);
        length += c * qpos[dex_ldg(m.jnt_qposadr + j)];
        vel += c * qvel[dex_ldg(m.jnt_dofadr + j)];
      }
    }
    length *= gear; vel *= gear;
    T c = st.ctrl[(size_t)envr * st.ctrl_stride + a];
    if (dex_ldg(m.act_ctrllimited + a)) c = dx_clamp(c, dex_ldg(m.act_ctrlrange + 2 * a), dex_ldg(m.act_ctrlrange + 2 * a + 1));
    T f = dex_ldg(m.act_gainprm + 3 * a) * c + dex_ldg(m.act_biasprm + 3 * a) + dex_ldg(m.act_biasprm + 3 * a + 1) * length +
          dex_ldg(m.act_biasprm + 3 * a + 2) * vel;
    if (dex_ldg(m.act_forcelimited + a)) f = dx_clamp(f, dex_ldg(m.act_forcerange + 2 * a), dex_ldg(m.act_forcerange + 2 * a + 1));
    if (joint) dex_atomic_add(act + dex_ldg(m.jnt_dofadr + tid), gear * f);
    else {
      const int wa = dex_ldg(m.ten_adr + tid), wn = dex_ldg(m.ten_num + tid);
      for (int w = wa; w < wa + wn; ++w)
        dex_atomic_add(act + dex_ldg(m.jnt_dofadr + dex_ldg(m.wrap_jnt + w)), gear * dex_ldg(m.wrap_coef + w) * f);
    }
  }
  // ---- per-tree reference points (lane = tree): FK along the path world -> reference body
  for (int t = 1 + gl; t < m.ntree; t += DEX_K0_G) {
    int path[DEX_MAXDEPTH], n = 0;
    for (int b = dex_ldg(m.tree_refbody + t); b > 0; b = dex_ldg(m.body_parent + b)) path[n++] = b;
    T pos[3] = {0, 0, 0}, quat[4] = {1, 0, 0, 0}, mat[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = n - 1; i >= 0; --i) {
      T npos[3], nquat[4], nmat[9];
      dex_body_fk<T>(m, qpos, path[i], dex_ldg(m.body_jntadr + path[i]), dex_ldg(m.body_jntnum + path[i]), 0, pos, quat, mat, npos,
                     nquat, nmat, nullptr, nullptr);
      for (int k = 0; k < 3; ++k) pos[k] = npos[k];
      for (int k = 0; k < 4; ++k) quat[k] = nquat[k];
      for (int k = 0; k < 9; ++k) mat[k] = nmat[k];
    }
    ref[3 * t] = pos[0]; ref[3 * t + 1] = pos[1]; ref[3 * t + 2] = pos[2];
  }
  dex_syncwarp();
  if (live && gl == 0) {  // world body record
    for (int k = 0; k < 16; ++k) xkin[k] = (k == 3 || k == 7 || k == 11 || k == 15) ? T(1) : T(0);
  }

  // ---- forward rounds (one per tree depth): lane state = pose / velocity / acceleration of its latest body
  T pos[3] = {0, 0, 0}, quat[4] = {1, 0, 0, 0}, mat[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  T vel[6] = {0, 0, 0, 0, 0, 0}, acc[6] = {0, 0, 0, 0, 0, 0};
  const int nround = m.k0_rounds;
  int recn[4];   // the next round's schedule record is requested one round ahead
  dex_ldg_int4(m.k0_rec + gl * 8, recn);
  for (int r = 0; r < nround; ++r) {
    // schedule record of this (round, lane): body, parent's lane, joints (adr | num << 16), tree id
    const int rec[4] = {recn[0], recn[1], recn[2], recn[3]};
    if (r + 1 < nround) dex_ldg_int4(m.k0_rec + ((r + 1) * DEX_K0_G + gl) * 8, recn);
    const int b = rec[0], src = rec[1], ja = rec[2] & 0xffff, jn = rec[2] >> 16, btree = rec[3];
    const int sl = (lane & ~(DEX_K0_G - 1)) | (src < 0 ? gl : src);
    T ppos[3], pquat[4], pmat[9], v[6], a[6];
    for (int k = 0; k < 3; ++k) ppos[k] = dex_shfl(pos[k], sl);
    for (int k = 0; k < 4; ++k) pquat[k] = dex_shfl(quat[k], sl);
    for (int k = 0; k < 9; ++k) pmat[k] = dex_shfl(mat[k], sl);
    for (int k = 0; k < 6; ++k) { v[k] = dex_shfl(vel[k], sl); a[k] = dex_shfl(acc[k], sl); }
    if (src < 0) {  // child of the world body
      for (int k = 0; k < 3; ++k) ppos[k] = 0;
      pquat[0] = 1; pquat[1] = pquat[2] = pquat[3] = 0;
      for (int k = 0; k < 9; ++k) pmat[k] = (k % 4 == 0) ? T(1) : T(0);
      for (int k = 0; k < 6; ++k) { v[k] = 0; a[k] = 0; }
      a[3] = -m.gravity[0]; a[4] = -m.gravity[1]; a[5] = -m.gravity[2];
    }
    if (b >= 0) {
      // packed body record: pos mass | quat | ipos - | inertia6 xfrc6, requested before the FK so that the loads overlap it
      // (K0 waits on loads half of the time, r02g profile; tree 0.140 -> 0.132 ms together with the ancestor lists below)
      T br[6][4];
#pragma unroll
      for (int k = 0; k < 6; ++k) dex_ldg4(m.body_rec + DEX_BODY_REC * b + 4 * k, br[k]);
      dex_body_fk<T>(m, qpos, b, ja, jn, btree, ppos, pquat, pmat, pos, quat, mat, ref, cdof);
      if (live) {
        const T rec[16] = {pos[0], pos[1], pos[2], quat[0], quat[1], quat[2], quat[3], mat[0], mat[1], mat[2], mat[3], mat[4],
                           mat[5], mat[6], mat[7], mat[8]};
        dex_store16(xkin + b * DEX_XKIN_BODY, rec);
      }
      // cinert about the tree reference point (mj_comPos)
      const T* rp = ref + 3 * btree;
      const T mass = br[0][3];
      const T ip[3] = {br[2][0], br[2][1], br[2][2]};
      T off[3], c[10];
      rot_vec(off, mat, ip);
      off[0] += pos[0] - rp[0]; off[1] += pos[1] - rp[1]; off[2] += pos[2] - rp[2];
      {
        // inertia6 = xx yy zz xy xz yz at record words 12..17
        const T Ib[9] = {br[3][0], br[3][3], br[4][0], br[3][3], br[3][1], br[4][1], br[4][0], br[4][1], br[3][2]};
        T RI[9];
        for (int rr = 0; rr < 3; ++rr) for (int cc = 0; cc < 3; ++cc)
          RI[3 * rr + cc] = mat[3 * rr] * Ib[cc] + mat[3 * rr + 1] * Ib[3 + cc] + mat[3 * rr + 2] * Ib[6 + cc];
        const T o2 = dot3(off, off);
        c[0] = RI[0] * mat[0] + RI[1] * mat[1] + RI[2] * mat[2] + mass * (o2 - off[0] * off[0]);
        c[1] = RI[3] * mat[3] + RI[4] * mat[4] + RI[5] * mat[5] + mass * (o2 - off[1] * off[1]);
        c[2] = RI[6] * mat[6] + RI[7] * mat[7] + RI[8] * mat[8] + mass * (o2 - off[2] * off[2]);
        c[3] = RI[0] * mat[3] + RI[1] * mat[4] + RI[2] * mat[5] - mass * off[0] * off[1];
        c[4] = RI[0] * mat[6] + RI[1] * mat[7] + RI[2] * mat[8] - mass * off[0] * off[2];
        c[5] = RI[3] * mat[6] + RI[4] * mat[7] + RI[5] * mat[8] - mass * off[1] * off[2];
        c[6] = mass * off[0]; c[7] = mass * off[1]; c[8] = mass * off[2]; c[9] = mass;
      }
      // velocity / acceleration through the body's joints (mj_comVel, mj_rne forward)
      {
        for (int j = ja; j < ja + jn; ++j) {
          int ji[4];
          dex_ldg_int4(m.jnt_irec + 4 * j, ji);
          const int da = ji[2];
          if (ji[0] == DEX_JNT_FREE) {
            for (int i = 0; i < 3; ++i) for (int k = 0; k < 6; ++k) v[k] += cdof[da + i][k] * qvel[da + i];
            T cdd[3][6];
            for (int i = 0; i < 3; ++i) cross_motion(cdd[i], v, cdof[da + 3 + i]);
            for (int i = 0; i < 3; ++i) for (int k = 0; k < 6; ++k) {
              v[k] += cdof[da + 3 + i][k] * qvel[da + 3 + i];
              a[k] += cdd[i][k] * qvel[da + 3 + i];
            }
          } else {
            T cdd[6];
            cross_motion(cdd, v, cdof[da]);
            for (int k = 0; k < 6; ++k) { v[k] += cdof[da][k] * qvel[da]; a[k] += cdd[k] * qvel[da]; }
          }
        }
      }
      for (int k = 0; k < 6; ++k) { vel[k] = v[k]; acc[k] = a[k]; }
      // own spatial force: I a + v x* (I v) - external force about the reference point
      {
        // xfrc6 = force(3) torque(3) at record words 18..23
        const T f[3] = {br[4][2], br[4][3], br[5][0]};
        T tq[3], t1[6], t2[6], t3[6];
        cross3(tq, off, f);
        mul_inert_vec(t1, c, a);
        mul_inert_vec(t2, c, v);
        cross_force(t3, v, t2);
        T* o = own + b * 16;
        for (int k = 0; k < 10; ++k) o[k] = c[k];
        o[10] = -(br[5][1] + tq[0]) + t1[0] + t3[0];
        o[11] = -(br[5][2] + tq[1]) + t1[1] + t3[1];
        o[12] = -(br[5][3] + tq[2]) + t1[2] + t3[2];
        o[13] = -f[0] + t1[3] + t3[3]; o[14] = -f[1] + t1[4] + t3[4]; o[15] = -f[2] + t1[5] + t3[5];
      }
    }
    dex_syncwarp();
  }
  // ---- backward rounds: a body's slot holds its subtree totals once all deeper rounds are done
  dex_ldg_int4(m.k0_rec + ((nround - 1) * DEX_K0_G + gl) * 8 + 4, recn);
  for (int r = nround - 1; r >= 0; --r) {
    const int rec[4] = {recn[0], recn[1], recn[2], recn[3]};   // body, dofs (adr | num << 16), children (adr | num << 16)
    if (r > 0) dex_ldg_int4(m.k0_rec + ((r - 1) * DEX_K0_G + gl) * 8 + 4, recn);
    const int b = rec[0];
    if (b >= 0) {
      // subtree totals = own + the totals the children left in their slots one round ago (no atomics: the r01k
      // profile had 22 % of K0's stall samples on shared-memory atomic adds into the palm's slot)
      int anc0[8];   // ancestor chain of the body's first dof, requested before the subtree sums
      { const int d0 = rec[1] & 0xffff; dex_ldg_int4(m.dof_anc8 + 8 * d0, anc0); dex_ldg_int4(m.dof_anc8 + 8 * d0 + 4, anc0 + 4); }
      T tot[16];
      for (int k = 0; k < 16; ++k) tot[k] = own[b * 16 + k];
      {
        const int ca = rec[2] & 0xffff, cn = rec[2] >> 16;
        for (int c = ca; c < ca + cn; ++c) {
          const T* oc = own + dex_ldg(m.body_child + c) * 16;
          for (int k = 0; k < 16; ++k) tot[k] += oc[k];
        }
        if (cn > 0) for (int k = 0; k < 16; ++k) own[b * 16 + k] = tot[k];
      }
      const int da = rec[1] & 0xffff, dn = rec[1] >> 16;
      for (int i = da; i < da + dn; ++i) {
        const T* ci = cdof[i];
        const T bias = ci[0] * tot[10] + ci[1] * tot[11] + ci[2] * tot[12] + ci[3] * tot[13] + ci[4] * tot[14] + ci[5] * tot[15];
        T buf[6];
        mul_inert_vec(buf, tot, ci);
        if (live) {
          tree[DEX_TREE_ROW_FS * 32 + i] = -dex_ldg(m.dof_damping + i) * qvel[i] - bias + act[i];  // qfrc_smooth
          for (int k = 0; k < 6; ++k) tree[k * 32 + i] = ci[k];
        }
        int anc[8];   // the dof and its ancestors (-1 terminated): two vector loads instead of a pointer chase
        if (i == da) { for (int t = 0; t < 8; ++t) anc[t] = anc0[t]; }
        else {
        dex_ldg_int4(m.dof_anc8 + 8 * i, anc);
        dex_ldg_int4(m.dof_anc8 + 8 * i + 4, anc + 4);
        }
        const T arm = dex_ldg(m.dof_armature + i);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int j = anc[t];
          if (j < 0) break;
          const T* cj = cdof[j];
          T s = cj[0] * buf[0] + cj[1] * buf[1] + cj[2] * buf[2] + cj[3] * buf[3] + cj[4] * buf[4] + cj[5] * buf[5];
          if (t == 0) s += arm;
          if (live) { M[i * 32 + j] = s; M[j * 32 + i] = s; }
        }
      }
    }
    dex_syncwarp();
  }
  if (live) for (int k = gl; k < 16; k += DEX_K0_G) tree[DEX_TREE_ROW_REF * 32 + k] = ref[k];
}

#endif  // DEX_TREE_CUH_
