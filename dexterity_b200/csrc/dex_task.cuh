// dex_task.cuh — fused pre-/post-step task kernels: the small
// per-step Python of the reference's composer hooks re-expressed as warp-per-env device code.
//
//   pre  : dm_env auto-reset of envs whose previous TimeStep was LAST (episode initialisation:
//          reference reorient.py:182-188, task.py:137-152), goal switching (task.py:154-165),
//          action -> ctrl (effectors/mujoco_actuation.py:30-33; clipping is left to the actuator
//          model exactly as upstream).
//   post : position/velocity refresh at the new state (dm_control legacy step: mj_step1 after
//          mj_step2), goal distance (goals/prop_orientation.py:40-50), success / timeout
//          bookkeeping (task.py:167-185), fall detection (reorient.py:229-235 via
//          utils/mujoco_collisions.py:95-119), reward (reorient.py:238-284, shared/rewards.py:13-15),
//          discount / termination (task.py:187-204, reorient.py:211-225), observations
//          (models/hands/dexterous_hand.py:254-310, primitive prop sensors), episode statistics.
#ifndef DEX_TASK_CUH_
#define DEX_TASK_CUH_

#include "../../include/dexterity_b200.h"
#include "dex_kin.cuh"

enum { DEX_TASK_REORIENT = 0, DEX_TASK_REACH = 1 };
enum { DEX_STEP_FIRST = 0, DEX_STEP_MID = 1, DEX_STEP_LAST = 2 };
// task_i columns
enum { TI_SUCCESSES = 0, TI_CHANGE_COUNTER = 1, TI_REGISTERED = 2, TI_EXCEEDED = 3, TI_GOAL_CHANGED = 4,
       TI_FAILURE = 5, TI_STEP_TYPE = 6, TI_EPISODE_LEN = 7, TI_COUNT = 8 };
// task_f columns
enum { TF_SOLVE_START = 0, TF_RETURN = 1, TF_GOAL_DIST = 2 /* up to 5 */, TF_COUNT = 8 };
// episode statistics (global, doubles)
enum { ES_RETURN = 0, ES_LENGTH = 1, ES_EPISODES = 2, ES_SUCCESS = 3, ES_FALL = 4, ES_TIMEOUT = 5, ES_PHYSICS_ERROR = 6,
       ES_COUNT = 8 };

#define DEX_OBS_STRIDE 128
#define DEX_GOAL_STRIDE 16

template <typename T>
struct TaskState {
  const T* action; int action_stride;   // [nenv][nu]
  T* obs;                               // [nenv][DEX_OBS_STRIDE]
  T* reward;                            // [nenv]
  T* discount;                          // [nenv]
  T* goal;                              // [nenv][DEX_GOAL_STRIDE]
  int* task_i;                          // [nenv][TI_COUNT]
  T* task_f;                            // [nenv][TF_COUNT]
  unsigned* rng;                        // [nenv][4]: key0 key1 counter unused
  unsigned char* skip;                  // [nenv] 1 = env was reset in this call (physics skipped)
  unsigned char* reset_mask;            // optional [nenv]: reset requests, consumed (cleared) by the pre-step
  double* stats;                        // [ES_COUNT]
};

// ---- Philox4x32-10 counter RNG (per-env stream keyed by seed and env id) --------------------
DEX_DEV void dex_philox(unsigned c[4], unsigned k0, unsigned k1) {
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c[0];
    const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c[2];
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c[1] ^ k0, n1 = (unsigned)p1;
    const unsigned n2 = (unsigned)(p0 >> 32) ^ c[3] ^ k1, n3 = (unsigned)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
template <typename T> DEX_DEV T dex_u01(unsigned x) { return (T)(x >> 8) * T(1.0 / 16777216.0); }
// 4 uniforms in [0,1) for (env, counter); all lanes compute the same values
template <typename T>
DEX_DEV void dex_rand4(const unsigned* rng, int env, unsigned counter, T* u) {
  unsigned c[4] = {counter, (unsigned)env, 0x5eedu, 0u};
  dex_philox(c, rng[0], rng[1]);
  for (int k = 0; k < 4; ++k) u[k] = dex_u01<T>(c[k]);
}
// dm_control rotations.UniformQuaternion (Shoemake)
template <typename T> DEX_DEV void dex_uniform_quat(const T* u, T* q) {
  const T tp = T(6.283185307179586);
  const T a = dx_sqrt(1 - u[0]), b = dx_sqrt(u[0]);
  q[0] = a * sin(tp * u[1]); q[1] = a * cos(tp * u[1]); q[2] = b * sin(tp * u[2]); q[3] = b * cos(tp * u[2]);
}

// ------------------------------------------------------------------------------------------
// pre-step
// ------------------------------------------------------------------------------------------
template <typename T>
DEX_DEV void dex_task_pre_env(const DevModel<T>& m, const BatchState<T>& st, const TaskState<T>& ts,
                              const DexTaskParams& p, int env, int force_reset) {
  const int lane = dex_lane();
  int* ti = ts.task_i + (size_t)env * TI_COUNT;
  T* tf = ts.task_f + (size_t)env * TF_COUNT;
  unsigned* rng = ts.rng + (size_t)env * 4;
  T* goal = ts.goal + (size_t)env * DEX_GOAL_STRIDE;
  const bool reset = force_reset || ti[TI_STEP_TYPE] == DEX_STEP_LAST || (ts.reset_mask && ts.reset_mask[env]);
  const unsigned ctr = rng[2];
  dex_syncwarp();
  if (reset) {
    T u0[4], u1[4], u2[4], gq[4], pq[4];
    dex_rand4<T>(rng, env, ctr, u0);
    dex_rand4<T>(rng, env, ctr + 1, u1);
    dex_rand4<T>(rng, env, ctr + 2, u2);
    dex_uniform_quat(u0, gq);       // first goal (task.py:145)
    dex_uniform_quat(u1, pq);       // prop orientation (reorient.py:149)
    for (int i = lane; i < m.nq; i += 32) {
      T v = dex_ldg(m.qpos0 + i);
      if (p.task == DEX_TASK_REORIENT) {
        const int k = i - p.prop_qposadr;
        if (k >= 0 && k < 3) v = (T)p.spawn_lo[k] + ((T)p.spawn_hi[k] - (T)p.spawn_lo[k]) * u2[k];
        else if (k >= 3 && k < 7) v = pq[k - 3];
      }
      st.qpos[(size_t)env * st.qpos_stride + i] = v;
    }
    if (lane < m.nv) { st.qvel[(size_t)env * st.qvel_stride + lane] = 0; st.warm[(size_t)env * st.warm_stride + lane] = 0; }
    if (lane < m.nu) st.ctrl[(size_t)env * st.ctrl_stride + lane] = 0;
    if (lane < 4 && p.task == DEX_TASK_REORIENT) goal[lane] = gq[lane];
    if (lane == 0) {
      if (ts.reset_mask) ts.reset_mask[env] = 0;   // the request is consumed: the env steps normally from the next call on
      st.time[env] = 0; st.status[env] = 0;
      if (st.time64) { st.time64[2 * (size_t)env] = 0; st.time64[2 * (size_t)env + 1] = 0; }
      ti[TI_SUCCESSES] = 0; ti[TI_CHANGE_COUNTER] = 0; ti[TI_REGISTERED] = 0; ti[TI_EXCEEDED] = 0;
      ti[TI_GOAL_CHANGED] = 1; ti[TI_FAILURE] = 0; ti[TI_EPISODE_LEN] = 0;
      tf[TF_SOLVE_START] = 0; tf[TF_RETURN] = 0;
      rng[2] = ctr + 3;
      ts.skip[env] = (p.task == DEX_TASK_REORIENT && p.place_max_attempts > 0) ? 2 : 1;   // 2: pose still to be collision-checked
    }
  } else {
    // GoalTask.before_step (task.py:154-165)
    const bool change = ti[TI_CHANGE_COUNTER] > p.steps_before_changing_goal;
    dex_syncwarp();
    if (change) {
      T u0[4], gq[4];
      dex_rand4<T>(rng, env, ctr, u0);
      dex_uniform_quat(u0, gq);
      if (lane < 4 && p.task == DEX_TASK_REORIENT) goal[lane] = gq[lane];
    }
    if (lane == 0) {
      if (change) {
        ti[TI_CHANGE_COUNTER] = 0; ti[TI_EXCEEDED] = 0; ti[TI_REGISTERED] = 0; ti[TI_GOAL_CHANGED] = 1;
        tf[TF_SOLVE_START] = st.time[env];
        if (st.time64) st.time64[2 * (size_t)env + 1] = st.time64[2 * (size_t)env];
        rng[2] = ctr + 1;
      } else ti[TI_GOAL_CHANGED] = 0;
      ts.skip[env] = 0;
    }
    // MujocoEffector.set_control: ctrl <- command (unclipped, as upstream stores it)
    if (lane < m.nu) st.ctrl[(size_t)env * st.ctrl_stride + lane] = ts.action[(size_t)env * ts.action_stride + lane];
  }
}

// ------------------------------------------------------------------------------------------
// Collision-checked prop placement (dm_control initializers.PropPlacer as configured in reference
// reorient.py:143-151: ignore_collisions=False, settle_physics=False, max_attempts_per_prop=20).
// One rejection round for an env whose skip byte is >= 2: the collision kernels have just run on the
// sampled pose (BatchState::skip_mode = 1); if the prop touches anything (a contact with dist <= 0
// that involves one of its geoms, or a full contact buffer, which upstream treats as a rejected
// pose too) a new position and orientation are drawn, else the env is marked placed (skip = 1).
// After place_max_attempts rejected poses PropPlacer raises; here the env gets DEX_STATUS_PLACEMENT
// and the post-step ends its episode at once, so that the next call draws again.
// ------------------------------------------------------------------------------------------
template <typename T>
DEX_DEV void dex_place_env(const DevModel<T>& m, const BatchState<T>& st, const TaskState<T>& ts,
                           const DexTaskParams& p, int env) {
  const int s = ts.skip[env];   // warp-uniform
  if (s < 2) return;
  const int lane = dex_lane();
  int ncon = st.ncon[env];
  if (ncon > DEX_MAXCON) ncon = DEX_MAXCON;
  const T* cin = st.contact + (size_t)env * DEX_MAXCON * DEX_CON_STRIDE;
  bool hit = (st.status[env] & (DEX_STATUS_CON_OVERFLOW | DEX_STATUS_CAND_OVERFLOW)) != 0;
  for (int c = lane; c < ncon; c += 32) {
    const T* cc = cin + c * DEX_CON_STRIDE;
    const int pr = (int)cc[7];
    const int b1 = dex_ldg(m.geom_body + dex_ldg(m.pair_geom1 + pr)), b2 = dex_ldg(m.geom_body + dex_ldg(m.pair_geom2 + pr));
    hit = hit || ((b1 == p.prop_body || b2 == p.prop_body) && cc[0] <= 0);
  }
  hit = dex_ballot(hit) != 0;
  unsigned* rng = ts.rng + (size_t)env * 4;
  const unsigned ctr = rng[2];
  dex_syncwarp();
  if (!hit) { if (lane == 0) ts.skip[env] = 1; return; }
  if (s - 1 >= p.place_max_attempts) {
    if (lane == 0) { st.status[env] |= DEX_STATUS_PLACEMENT; ts.skip[env] = 1; }
    return;
  }
  T u1[4], u2[4], pq[4];
  dex_rand4<T>(rng, env, ctr, u1);
  dex_rand4<T>(rng, env, ctr + 1, u2);
  dex_uniform_quat(u1, pq);
  T* qp = st.qpos + (size_t)env * st.qpos_stride + p.prop_qposadr;
  if (lane < 3) qp[lane] = (T)p.spawn_lo[lane] + ((T)p.spawn_hi[lane] - (T)p.spawn_lo[lane]) * u2[lane];
  else if (lane < 7) qp[lane] = pq[lane - 3];
  if (lane == 0) {
    rng[2] = ctr + 2;
    st.status[env] &= ~(DEX_STATUS_CON_OVERFLOW | DEX_STATUS_CAND_OVERFLOW);   // they described the rejected pose
    ts.skip[env] = (unsigned char)(s + 1);
  }
}

// ------------------------------------------------------------------------------------------
// post-step.  smem (units of T): xpos | xquat | xmat | xanchor | xaxis | ref | qpos | qvel
// ------------------------------------------------------------------------------------------
#define DEX_K3_SMEM_T (DEX_MAXBODY * 16 + DEX_MAXJNT * 6 + 16 + 48 + 32)

template <typename T>
DEX_DEV void dex_task_post_env(const DevModel<T>& m, const BatchState<T>& st, const TaskState<T>& ts,
                               const DexTaskParams& p, int env, T* smem) {
  const int lane = dex_lane();
  T* xpos = smem;
  T* xquat = xpos + DEX_MAXBODY * 3;
  T* xmat = xquat + DEX_MAXBODY * 4;
  T* xanchor = xmat + DEX_MAXBODY * 9;
  T* xaxis = xanchor + DEX_MAXJNT * 3;
  T* ref = xaxis + DEX_MAXJNT * 3;
  T* qpos = ref + 16;
  T* qvel = qpos + 48;
  int* ti = ts.task_i + (size_t)env * TI_COUNT;
  T* tf = ts.task_f + (size_t)env * TF_COUNT;
  T* obs = ts.obs + (size_t)env * DEX_OBS_STRIDE;
  const T* goal = ts.goal + (size_t)env * DEX_GOAL_STRIDE;
  const int nv = m.nv;

  for (int i = lane; i < m.nq; i += 32) qpos[i] = st.qpos[(size_t)env * st.qpos_stride + i];
  const T qv = lane < nv ? st.qvel[(size_t)env * st.qvel_stride + lane] : T(0);
  qvel[lane] = qv;
  const T ctrl = lane < m.nu ? st.ctrl[(size_t)env * st.ctrl_stride + lane] : T(0);
  const bool fresh = ts.skip[env] != 0;
  dex_syncwarp();
  KinBuf<T> kb = {xpos, xquat, xmat, xanchor, xaxis};
  dex_fk(m, qpos, kb);
  if (lane < m.ntree) {
    const int rb = dex_ldg(m.tree_refbody + lane);
    ref[3 * lane] = xpos[3 * rb]; ref[3 * lane + 1] = xpos[3 * rb + 1]; ref[3 * lane + 2] = xpos[3 * rb + 2];
  }
  dex_syncwarp();
  T cd[6];
  dex_cdof_lane(m, lane, xpos, xmat, xanchor, xaxis, ref, cd);

  // ---- observations shared by both tasks: hand joints, fingertip sites
  if (lane < p.nhand) {
    const T q = qpos[lane];
    obs[2 * lane] = sin(q); obs[2 * lane + 1] = cos(q);        // dexterous_hand.py:254-260 (interleaved)
    obs[2 * p.nhand + lane] = qv;                              // :262-264
  }
  const int o_tip = 3 * p.nhand, o_tipv = o_tip + 3 * p.nsite;
  T tip[5][3];
  for (int s = 0; s < p.nsite; ++s) {
    const int sid = p.site_id[s], b = dex_ldg(m.site_body + sid);
    T sp[3] = {dex_ldg(m.site_pos + 3 * sid), dex_ldg(m.site_pos + 3 * sid + 1), dex_ldg(m.site_pos + 3 * sid + 2)}, w[3];
    rot_vec(w, xmat + 9 * b, sp);
    w[0] += xpos[3 * b]; w[1] += xpos[3 * b + 1]; w[2] += xpos[3 * b + 2];
    tip[s][0] = w[0]; tip[s][1] = w[1]; tip[s][2] = w[2];
    // world-frame linear velocity of the site: mj_objectVelocity(flg_local=0)[3:6] (mujoco_utils.py:10-35)
    T v[3] = {0, 0, 0};
    if (lane < nv && (((unsigned)dex_ldg(m.body_dofmask + b) >> lane) & 1u)) {
      const T* rp = ref + 3 * dex_ldg(m.dof_tree + lane);
      T off[3] = {w[0] - rp[0], w[1] - rp[1], w[2] - rp[2]};
      cross3(v, cd, off);
      v[0] = (v[0] + cd[3]) * qv; v[1] = (v[1] + cd[4]) * qv; v[2] = (v[2] + cd[5]) * qv;
    }
    v[0] = dex_warp_sum(v[0]); v[1] = dex_warp_sum(v[1]); v[2] = dex_warp_sum(v[2]);
    if (lane < 3) { obs[o_tip + 3 * s + lane] = w[lane]; obs[o_tipv + 3 * s + lane] = v[lane]; }
  }
  int o = o_tipv + 3 * p.nsite;

  // ---- task specific: goal distance, fall, extra observations
  T dist[5] = {0, 0, 0, 0, 0};
  int ndist = 1;
  bool fallen = false;
  if (p.task == DEX_TASK_REORIENT) {
    const int qa = p.prop_qposadr, da = p.prop_dofadr, pb = p.prop_body;
    const T* cq = xquat + 4 * pb;  // framequat of the prop geom (goals/prop_orientation.py:31-32)
    // quat_diff_active(source=current, target=goal).w = <goal, current>; axis-angle norm wrapped to [0, pi]
    T w = goal[0] * cq[0] + goal[1] * cq[1] + goal[2] * cq[2] + goal[3] * cq[3];
    w = dx_clamp(w, T(-1), T(1));
    T ang = 2 * acos(w);
    const T pi = T(3.141592653589793);
    if (ang > pi) ang = 2 * pi - ang;
    dist[0] = ang;
    // fall: prop-ground contact at the refreshed state (mjc_PlaneBox, dist <= margin 1e-8)
    if (p.fall_termination) {
      const int gg = p.ground_geom, pg = p.prop_geom;
      const int gbody = dex_ldg(m.geom_body + gg);
      T nl[3] = {dex_ldg(m.geom_mat + 9 * gg + 2), dex_ldg(m.geom_mat + 9 * gg + 5), dex_ldg(m.geom_mat + 9 * gg + 8)}, n[3], pp[3], gp[3];
      rot_vec(n, xmat + 9 * gbody, nl);
      T gl[3] = {dex_ldg(m.geom_pack + 12 * gg), dex_ldg(m.geom_pack + 12 * gg + 1), dex_ldg(m.geom_pack + 12 * gg + 2)};
      rot_vec(gp, xmat + 9 * gbody, gl);
      gp[0] += xpos[3 * gbody]; gp[1] += xpos[3 * gbody + 1]; gp[2] += xpos[3 * gbody + 2];
      T pl[3] = {dex_ldg(m.geom_pack + 12 * pg), dex_ldg(m.geom_pack + 12 * pg + 1), dex_ldg(m.geom_pack + 12 * pg + 2)};
      rot_vec(pp, xmat + 9 * pb, pl);
      T dif[3] = {pp[0] + xpos[3 * pb] - gp[0], pp[1] + xpos[3 * pb + 1] - gp[1], pp[2] + xpos[3 * pb + 2] - gp[2]};
      const T dc = dot3(dif, n);
      bool hit = false;
      if (lane < 8) {
        T sz[3] = {dex_ldg(m.geom_size + 3 * pg), dex_ldg(m.geom_size + 3 * pg + 1), dex_ldg(m.geom_size + 3 * pg + 2)};
        T vec[3] = {(lane & 1) ? sz[0] : -sz[0], (lane & 2) ? sz[1] : -sz[1], (lane & 4) ? sz[2] : -sz[2]}, G[9], t1[3], c[3];
        for (int i = 0; i < 9; ++i) G[i] = dex_ldg(m.geom_mat + 9 * pg + i);
        rot_vec(t1, G, vec);
        rot_vec(c, xmat + 9 * pb, t1);
        const T ld = dot3(n, c);
        hit = !(dc + ld > 0 || ld > 0);
      }
      fallen = dex_ballot(hit) != 0;
    }
    // prop / target observables (primitive.Primitive sensors enabled by reorient.py:81-86)
    if (lane < 3) {
      obs[o + lane] = xpos[3 * pb + lane];                                  // prop/position
      obs[o + 7 + lane] = qvel[da + lane];                                  // prop/linear_velocity (world)
      const T* R = xmat + 9 * pb;
      obs[o + 10 + lane] = R[3 * lane] * qvel[da + 3] + R[3 * lane + 1] * qvel[da + 4] + R[3 * lane + 2] * qvel[da + 5];  // angular (world)
    }
    if (lane < 4) {
      obs[o + 3 + lane] = cq[lane];                                         // prop/orientation
      obs[o + 13 + lane] = goal[lane];                                      // target_prop/orientation (hint pose = goal)
      obs[o + 17 + lane] = goal[lane];                                      // goal_state (task.py:206-216)
    }
    (void)qa;
  } else {
    ndist = p.nsite;
    for (int s = 0; s < p.nsite; ++s) {
      T d[3] = {goal[3 * s] - tip[s][0], goal[3 * s + 1] - tip[s][1], goal[3 * s + 2] - tip[s][2]};
      dist[s] = dx_sqrt(dot3(d, d));   // goals/fingertip_position.py:126-130
    }
    if (lane < 3 * p.nsite) obs[o + lane] = goal[lane];                      // goal_state
  }

  // ---- TimeStep assembly
  if (fresh) {
    if (lane == 0) {
      ts.reward[env] = 0; ts.discount[env] = 1; ti[TI_STEP_TYPE] = DEX_STEP_FIRST; ti[TI_FAILURE] = 0;
      for (int s = 0; s < ndist; ++s) tf[TF_GOAL_DIST + s] = dist[s];
    }
    return;
  }
  const T act2 = dex_warp_sum(ctrl * ctrl);  // ||physics.data.ctrl||^2 (reorient.py:277)
  if (lane == 0) {
    // data.time and _solve_start_time in double, as the reference holds them (task.py:149,161,184): in FP32 1500 additions
    // of 0.005 drift enough to move the 7.5 s per-goal timeout by a control step
    const double time = st.time64 ? st.time64[2 * (size_t)env] : (double)st.time[env];
    const double solve_start = st.time64 ? st.time64[2 * (size_t)env + 1] : (double)tf[TF_SOLVE_START];
    const unsigned status = st.status[env];
    bool all_in = true;
    for (int s = 0; s < ndist; ++s) { all_in = all_in && (dist[s] <= (T)p.success_threshold); tf[TF_GOAL_DIST + s] = dist[s]; }
    int successes = ti[TI_SUCCESSES], exceeded = ti[TI_EXCEEDED];
    if (all_in) {                                   // task.py:177-181
      ti[TI_CHANGE_COUNTER] += 1;
      if (!ti[TI_REGISTERED]) { successes += 1; ti[TI_REGISTERED] = 1; }
    } else if (p.has_max_time && time - solve_start > p.max_time_per_goal) exceeded = 1;  // :182-185
    ti[TI_SUCCESSES] = successes; ti[TI_EXCEEDED] = exceeded;
    const bool failure = p.task == DEX_TASK_REORIENT && p.fall_termination && fallen;
    ti[TI_FAILURE] = failure ? 1 : 0;
    const bool solved = successes >= p.successes_needed;
    const bool timeout = p.has_max_time && exceeded;
    const bool physics_error = (status & (DEX_STATUS_BADQACC | DEX_STATUS_PLACEMENT)) != 0;
    const bool over_time = time >= p.time_limit;
    const bool terminate = solved || timeout || failure || physics_error || over_time;
    T reward, discount;
    if (p.task == DEX_TASK_REORIENT) {
      const T d = dist[0];
      reward = (T)p.w_orientation / (d + (T)p.orientation_eps) +
               (T)p.w_success_bonus * ((d >= 0 && d <= (T)p.bonus_threshold) ? T(1) : T(0)) +
               (T)p.w_action_smoothing * act2;
      discount = failure ? T(1) : (solved ? T(0) : T(1));   // reorient.py:222-225, task.py:195-204
    } else {
      // reach.py:196-210: mean over fingers of 0 (inside 1 cm) else -tanh^2(w d) (dense) / -1 (sparse)
      T acc = 0;
      for (int s = 0; s < ndist; ++s) {
        if (dist[s] <= T(0.01)) continue;
        if (p.dense_reward) { const T tt = tanh((T)p.reach_tanh_weight * dist[s]); acc -= tt * tt; }
        else acc -= 1;
      }
      reward = acc / (T)ndist;
      discount = solved ? T(0) : T(1);
    }
    if (physics_error) { reward = 0; discount = 0; }   // dm_control PhysicsError handling
    ts.reward[env] = reward;
    ts.discount[env] = discount;
    ti[TI_STEP_TYPE] = terminate ? DEX_STEP_LAST : DEX_STEP_MID;
    ti[TI_EPISODE_LEN] += 1;
    tf[TF_RETURN] += reward;
    if (terminate && ts.stats) {
      dex_atomic_add(ts.stats + ES_RETURN, (double)tf[TF_RETURN]);
      dex_atomic_add(ts.stats + ES_LENGTH, (double)ti[TI_EPISODE_LEN]);
      dex_atomic_add(ts.stats + ES_EPISODES, 1.0);
      if (solved) dex_atomic_add(ts.stats + ES_SUCCESS, 1.0);
      if (failure) dex_atomic_add(ts.stats + ES_FALL, 1.0);
      if (timeout) dex_atomic_add(ts.stats + ES_TIMEOUT, 1.0);
      if (physics_error) dex_atomic_add(ts.stats + ES_PHYSICS_ERROR, 1.0);
    }
  }
}

#endif  // DEX_TASK_CUH_
