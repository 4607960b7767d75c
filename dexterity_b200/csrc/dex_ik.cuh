// dex_ik.cuh — K6: batched damped-least-squares fingertip IK, one warp per target set
// (SURVEY.md §8a row 22).  Restates reference inverse_kinematics/ik_solver.py:71-228 and
// controllers/dls/dls.py:43-77:
//   for each attempt (mid-range start, then uniform random restarts):
//     repeat <= max_steps: twist = 0.95 (target - x); J = stacked site jacp (3*nsite x nv);
//       qdot = (J^T J + lambda I)^-1 J^T twist; q += qdot (dt = 1); clip to joint range; FK;
//       stop on tolerance (early_stop) or "not enough progress" (err / step > 20)
//     accept if every fingertip is within linear_tol; keep the solution nearest the mid-range.
// Lane = dof.  The normal matrix is assembled row-per-lane and factorised with the same warp
// Cholesky as the Newton solver.
#ifndef DEX_IK_CUH_
#define DEX_IK_CUH_

#include "../../include/dexterity_b200.h"
#include "dex_dynamics.cuh"
#include "dex_task.cuh"

#define DEX_IK_MAXSITE 5
#define DEX_K6_OFF_XPOS 0
#define DEX_K6_OFF_XQUAT (DEX_K6_OFF_XPOS + DEX_MAXBODY * 3)
#define DEX_K6_OFF_XMAT (DEX_K6_OFF_XQUAT + DEX_MAXBODY * 4)
#define DEX_K6_OFF_XANCHOR (DEX_K6_OFF_XMAT + DEX_MAXBODY * 9)
#define DEX_K6_OFF_XAXIS (DEX_K6_OFF_XANCHOR + DEX_MAXJNT * 3)
#define DEX_K6_OFF_QPOS (DEX_K6_OFF_XAXIS + DEX_MAXJNT * 3)
#define DEX_K6_OFF_JP (DEX_K6_OFF_QPOS + 48)
#define DEX_K6_OFF_H (DEX_K6_OFF_JP + DEX_IK_MAXSITE * 32 * 4)
#define DEX_K6_SMEM_T (DEX_K6_OFF_H + 32 * DEX_MSTRIDE)

// site positions for the current qpos (FK) -> cur[s][3] (warp-uniform registers)
template <typename T>
DEX_DEV void dex_ik_sites(const DevModel<T>& m, const DexIkParams& p, const T* xpos, const T* xmat, T cur[][3]) {
  for (int s = 0; s < p.nsite; ++s) {
    const int sid = p.site_id[s], b = dex_ldg(m.site_body + sid);
    T sp[3] = {dex_ldg(m.site_pos + 3 * sid), dex_ldg(m.site_pos + 3 * sid + 1), dex_ldg(m.site_pos + 3 * sid + 2)}, w[3];
    rot_vec(w, xmat + 9 * b, sp);
    cur[s][0] = w[0] + xpos[3 * b]; cur[s][1] = w[1] + xpos[3 * b + 1]; cur[s][2] = w[2] + xpos[3 * b + 2];
  }
}

template <typename T>
DEX_DEV void dex_ik_env(const DevModel<T>& m, const DexIkParams& p, const T* targets, T* qout, int qout_stride,
                        unsigned char* ok_out, int* attempts_out, int idx, T* smem) {
  const int lane = dex_lane();
  const int nv = m.nv;
  T* xpos = smem + DEX_K6_OFF_XPOS;
  T* xquat = smem + DEX_K6_OFF_XQUAT;
  T* xmat = smem + DEX_K6_OFF_XMAT;
  T* xanchor = smem + DEX_K6_OFF_XANCHOR;
  T* xaxis = smem + DEX_K6_OFF_XAXIS;
  T* qpos = smem + DEX_K6_OFF_QPOS;
  T* jp = smem + DEX_K6_OFF_JP;   // [site][dof][4]
  T* H = smem + DEX_K6_OFF_H;
  KinBuf<T> kb = {xpos, xquat, xmat, xanchor, xaxis};
  const bool isdof = lane < nv;
  // hinge dofs only (validated on the host): dof i <-> joint i <-> qpos i
  const int jid = isdof ? dex_ldg(m.dof_jnt + lane) : 0;
  const int qadr = isdof ? dex_ldg(m.jnt_qposadr + jid) : 0;
  const T lo = isdof ? dex_ldg(m.jnt_range + 2 * jid) : T(0), hi = isdof ? dex_ldg(m.jnt_range + 2 * jid + 1) : T(0);
  const T mid = T(0.5) * (lo + hi);            // nullspace reference: joint mid-range (ik_solver.py:58)
  T target[DEX_IK_MAXSITE][3];
  for (int s = 0; s < p.nsite; ++s) for (int k = 0; k < 3; ++k) target[s][k] = targets[(size_t)idx * 3 * p.nsite + 3 * s + k];
  const unsigned rng[4] = {(unsigned)(p.seed & 0xffffffffu), (unsigned)(p.seed >> 32), 0u, 0u};

  bool success = false;
  T best_err = Real<T>::big(), best_q = mid;
  int attempt = 0;
  for (; attempt < p.num_attempts; ++attempt) {
    // ---- initial configuration (ik_solver.py:125-130)
    T q = mid;
    if (attempt > 0) {
      T u[4];
      dex_rand4<T>(rng, idx * 64 + attempt, (unsigned)(lane >> 2), u);
      q = lo + (hi - lo) * u[lane & 3];
    }
    dex_syncwarp();
    if (isdof) qpos[qadr] = q;
    dex_syncwarp();
    dex_fk(m, qpos, kb);
    T cur[DEX_IK_MAXSITE][3], prev[DEX_IK_MAXSITE][3], err[DEX_IK_MAXSITE];
    dex_ik_sites(m, p, xpos, xmat, cur);
    for (int s = 0; s < p.nsite; ++s) { for (int k = 0; k < 3; ++k) prev[s][k] = cur[s][k]; err[s] = Real<T>::big(); }
    for (int step = 0; step < p.max_steps; ++step) {
      // ---- positional Jacobian columns of this lane's dof for every site (mj_jacSite, jacp only)
      T rhs = 0;
      for (int s = 0; s < p.nsite; ++s) {
        const int b = dex_ldg(m.site_body + p.site_id[s]);
        T v[3] = {0, 0, 0};
        if (isdof && (((unsigned)dex_ldg(m.body_dofmask + b) >> lane) & 1u)) {
          T ax[3] = {xaxis[3 * jid], xaxis[3 * jid + 1], xaxis[3 * jid + 2]};
          T off[3] = {cur[s][0] - xanchor[3 * jid], cur[s][1] - xanchor[3 * jid + 1], cur[s][2] - xanchor[3 * jid + 2]};
          cross3(v, ax, off);
        }
        jp[(s * 32 + lane) * 4] = v[0]; jp[(s * 32 + lane) * 4 + 1] = v[1]; jp[(s * 32 + lane) * 4 + 2] = v[2];
        // twist = gain * (target - cur) / dt  (ik_solver.py:253-261), rhs = J^T twist
        rhs += (T)p.gain * (v[0] * (target[s][0] - cur[s][0]) + v[1] * (target[s][1] - cur[s][1]) + v[2] * (target[s][2] - cur[s][2]));
      }
      dex_syncwarp();
      // ---- normal matrix row: (J^T J + lambda I)[lane][j]  (dls.py:69-74)
      T hd = 1;
      for (int j = 0; j < 32; ++j) {
        T a = (j == lane) ? (T)p.regularization : T(0);
        for (int s = 0; s < p.nsite; ++s) {
          const T* x = jp + (s * 32 + lane) * 4;
          const T* y = jp + (s * 32 + j) * 4;
          a += x[0] * y[0] + x[1] * y[1] + x[2] * y[2];
        }
        const bool in = isdof && j < nv;
        if (in && j == lane) hd = a;
        H[lane * DEX_MSTRIDE + j] = (in && j < lane) ? a : T(0);
      }
      dex_syncwarp();
      int fail = 0;
      const T hinv = dex_chol_factor(H, hd, nv, nv, &fail);
      const T qdot = dex_chol_apply(H, nv, hinv, isdof ? rhs : T(0));
      // ---- integrate (dt = 1), clip to the joint range (ik_solver.py:189-194, 240-246)
      q = dx_clamp(q + qdot, lo, hi);
      dex_syncwarp();
      if (isdof) qpos[qadr] = q;
      dex_syncwarp();
      dex_fk(m, qpos, kb);
      dex_ik_sites(m, p, xpos, xmat, cur);
      bool close_enough = true, not_enough_progress = false;
      for (int s = 0; s < p.nsite; ++s) {
        T d[3] = {target[s][0] - cur[s][0], target[s][1] - cur[s][1], target[s][2] - cur[s][2]};
        T c[3] = {cur[s][0] - prev[s][0], cur[s][1] - prev[s][1], cur[s][2] - prev[s][2]};
        err[s] = dx_sqrt(dot3(d, d));
        const T change = dx_sqrt(dot3(c, c));
        if (err[s] > (T)p.linear_tol) close_enough = false;
        if (err[s] / (change + T(1e-10)) > (T)p.progress_threshold) not_enough_progress = true;
        for (int k = 0; k < 3; ++k) prev[s][k] = cur[s][k];
      }
      if ((p.early_stop && close_enough) || not_enough_progress) break;
    }
    bool all_ok = true;
    for (int s = 0; s < p.nsite; ++s) if (!(err[s] <= (T)p.linear_tol)) all_ok = false;
    if (all_ok) {   // ik_solver.py:138-145
      success = true;
      const T dq = isdof ? (q - mid) : T(0);
      const T ns = dx_sqrt(dex_warp_sum(dq * dq));
      if (ns < best_err) { best_err = ns; best_q = q; }
    }
    if (success && p.stop_on_first) { ++attempt; break; }
  }
  if (isdof) qout[(size_t)idx * qout_stride + qadr] = success ? best_q : mid;
  if (lane == 0) { ok_out[idx] = success ? 1 : 0; if (attempts_out) attempts_out[idx] = attempt; }
}

#endif  // DEX_IK_CUH_
