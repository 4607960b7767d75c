// dex_model.cuh — device-side model constants (one copy per GPU, shared by all envs) and
// small vector/quaternion helpers.  Populated by dex_api.cu from the host blob
// (include/dex_model_blob.h); arrays are converted to the batch scalar type T.
#ifndef DEX_MODEL_CUH_
#define DEX_MODEL_CUH_

#include "dex_portable.cuh"

#define DEX_MAXBODY 40
#define DEX_MAXDOF 32
#define DEX_MAXJNT 32
#define DEX_MAXU 32
#define DEX_MAXTEN 8
#ifndef DEX_MAXCON
#define DEX_MAXCON 32      // per-env contact capacity = one lane per contact in K2 (overflow sets a status bit)
#endif
#define DEX_MAXCAND 128    // per-env narrowphase candidate capacity
#define DEX_CON_STRIDE 8   // words per stored contact: dist, pos(3), normal(3), pair id
#define DEX_CUBEMAP_WORDS 192  // per mesh geom: 6 x 8 x 8 start vertices for the hill-climbing support, 2 per word

// tables of the sparse solver (dex_sparse.cuh).  Per elimination structure: the warp-uniform part travels in the kernel
// parameters (DexSpConst), the per-lane part in DevModel::sp_tab as [structure][DEX_SPL_WORDS][32 lanes].
#define DEX_SP_SLOTS 16
#define DEX_SPL_DEPTH 0   // slot of the lane's diagonal = number of its ancestors
#define DEX_SPL_STEP 1    // step at which the lane is eliminated (255: lane beyond nv)
#define DEX_SPL_ANC 2     // 4 words: ancestor lane per slot, one byte each (own lane as padding)
#define DEX_SPL_DESC 6    // 4 words: lane whose row this lane pulls at step s, one byte each (0xff: none)
#define DEX_SPL_SCHUR 10  // the lane's entries of the deferred trunk update: t | j << 4 | t2 << 8 | j2 << 12 (0xf: none)
#define DEX_SPL_WORDS 11
struct DexSpConst {
  int nstep, nslot;          // elimination steps; off-diagonal slots in use (= maximum depth)
  int nt;                    // trunk dofs (depths 0 .. nt-1)
  int first_trunk_step;      // the deferred update of the trunk rows runs before this step
  int ndef;                  // rows below the trunk
  unsigned char def[24];     // their lanes
  unsigned char trunk[8];    // lane of the trunk dof at depth d
};

// per-env status bits (SURVEY.md §5 failure detection: mirrors upstream PhysicsError handling)
#define DEX_STATUS_BADQACC 1u
#define DEX_STATUS_CON_OVERFLOW 2u
#define DEX_STATUS_CAND_OVERFLOW 4u
#define DEX_STATUS_SOLVER_MAXITER 8u
#define DEX_STATUS_CHOL_FAIL 16u
#define DEX_STATUS_PLACEMENT 32u   // no collision-free prop pose within DexTaskParams::place_max_attempts (dm_control PropPlacer raises)

enum { DEX_JNT_FREE = 0, DEX_JNT_BALL = 1, DEX_JNT_SLIDE = 2, DEX_JNT_HINGE = 3 };
enum { DEX_GEOM_PLANE = 0, DEX_GEOM_SPHERE = 2, DEX_GEOM_CAPSULE = 3, DEX_GEOM_BOX = 6, DEX_GEOM_MESH = 7 };

// X(type, name): type 'I' = int32 array, 'R' = real array
#define DEX_DEVMODEL_ARRAYS(X)                                                             \
  X(I, body_parent) X(I, body_depth) X(I, body_jntadr) X(I, body_jntnum) X(I, body_dofadr) \
  X(I, body_dofnum) X(I, body_tree) X(I, body_dofmask) X(I, body_weld)                     \
  X(R, body_pos) X(R, body_quat) X(R, body_ipos) X(R, body_mass) X(R, body_inertia6)       \
  X(R, body_xfrc) X(R, body_invweight0) X(R, body_bsphere) X(R, body_pack)                                 \
  X(R, body_rec) X(R, jnt_rec) X(I, jnt_irec) X(I, tree_refbody) X(I, k0_rec) X(I, dof_anc8) X(I, body_child)                                            \
  X(I, jnt_type) X(I, jnt_body) X(I, jnt_qposadr) X(I, jnt_dofadr) X(I, jnt_limited)       \
  X(R, jnt_pos) X(R, jnt_axis) X(R, jnt_range) X(R, jnt_margin) X(R, jnt_solref)           \
  X(R, jnt_solimp) X(R, qpos0)                                                             \
  X(I, dof_body) X(I, dof_jnt) X(I, dof_parent) X(I, dof_tree)                             \
  X(R, dof_armature) X(R, dof_damping) X(R, dof_frictionloss) X(R, dof_invweight0)         \
  X(R, dof_solref) X(R, dof_solimp)                                                        \
  X(I, ten_adr) X(I, ten_num) X(I, ten_limited) X(R, ten_range) X(R, ten_margin)           \
  X(R, ten_invweight0) X(R, ten_solref) X(R, ten_solimp) X(I, wrap_jnt) X(R, wrap_coef)    \
  X(I, act_trntype) X(I, act_trnid) X(I, act_ctrllimited) X(I, act_forcelimited)           \
  X(R, act_gear) X(R, act_gainprm) X(R, act_biasprm) X(R, act_ctrlrange) X(R, act_forcerange) \
  X(I, geom_type) X(I, geom_body) X(I, geom_vertadr) X(I, geom_vertnum)                    \
  X(R, geom_pos) X(R, geom_mat) X(R, geom_size) X(R, geom_rbound) X(R, geom_aabb) X(R, geom_pack) X(R, vert)               \
  X(I, geom_adjbase) X(I, geom_cubemap) X(I, vadj) X(I, adj)                                \
  X(I, bp_body1) X(I, bp_body2) X(I, bp_pairadr) X(I, bp_pairnum)                          \
  X(I, pair_geom1) X(I, pair_geom2) X(I, pair_gg) X(I, pair_condim) X(R, pair_friction) X(R, pair_solref) \
  X(R, pair_solimp) X(R, pair_margin) X(R, pair_gap)                                       \
  X(I, site_body) X(R, site_pos)                                                           \
  X(I, dof_pidx) X(R, dof_flrow) X(I, arrow_tab) X(I, sp_tab) X(I, body_augmask)

template <typename T>
struct DevModel {
  int nbody, njnt, nq, nv, nu, ntendon, ngeom, npair, nsite, nbodypair, ntree, maxdepth;
  int nconmax, iterations, mpr_iterations, nv_euler;
  int chol_split;  // first dof (multiple of 4) of a trailing group of kinematic trees that M does not couple to the dofs before it (nv if none)
  int k0_rounds;  // rounds of the 8-lanes-per-env tree kernel schedule (0 = model does not fit: thread-per-env kernel)
  T timestep, gravity[3], tolerance, mpr_tolerance, meaninertia, max_margin;
  double timestep_d;  // opt.timestep unrounded: data.time advances in double whatever the batch dtype
  // "arrowhead" decomposition of the dofs for the structured solver kernel (dex_arrow.cuh): up to 5 BRANCHES of up to 5 dofs
  // (subtrees that M does not couple to each other, e.g. the fingers) and a BORDER of up to 8 dofs (the dofs above the
  // branching point, e.g. the wrist, followed by the dofs of unbranched trees, e.g. a free object).  -1 = identity padding.
  int arrow_ok;          // 0: the model does not fit, every env takes the dense solver
  int arrow_nmo;         // leading border rows that M couples to the branches (<= 2); they are also the border rows with joint damping
  int arrow_bdof[5][5];  // dof of (branch, slot), root to leaf
  int arrow_odof[8];     // dof of border row r
  int arrow_bmask[5];    // dof bit mask of each branch
  int arrow_omask;       // dof bit mask of the border
  int arrow_rowdof[24];  // dofs that own friction-loss / limit rows (-1 padded), dealt three to a lane in the row view
  int sp_ok;             // sparse solver tables valid (dex_sparse.cuh)
  DexSpConst sp[2];      // the two elimination structures (0: dof forest, 1: chained)
  // the same tables as a device array for lane-indexed reads (arrow_tab): [8 lanes][8] = branch dofs 0-4, border dof 5, -, - then [24] rows
#define DEX_AR_TAB_ROWS 64
#define DEX_X_I(name) const int* name;
#define DEX_X_R(name) const T* name;
#define DEX_X(t, name) DEX_X_##t(name)
  DEX_DEVMODEL_ARRAYS(DEX_X)
#undef DEX_X
#undef DEX_X_I
#undef DEX_X_R
};

// Per-batch state pointers (torch-owned storage bound through dex_bind).
// Row-major [num_envs][stride]; strides are in elements.
template <typename T>
struct BatchState {
  int nenv;                       // one past the last env of this launch
  int env0;                       // first env of this launch (env chunks run as independent pipelines)
  T* qpos; int qpos_stride;
  T* qvel; int qvel_stride;
  T* ctrl; int ctrl_stride;
  T* warm; int warm_stride;       // qacc_warmstart
  T* time;                        // [nenv]
  unsigned* status;               // [nenv] DEX_STATUS_* bits (sticky until reset)
  // batch-owned scratch
  T* contact;                     // [nenv][DEX_MAXCON][DEX_CON_STRIDE]
  int* ncon;                      // [nenv]
  int* niter;                     // [nenv] solver iterations of the last mj_step
  const unsigned char* skip;      // optional [nenv]: 1 = leave this env untouched (it was just reset)
  T* jc;                          // [nenv][3*DEX_MAXCON][32] contact-frame Jacobian scratch (L2 resident)
  T* mscr;                        // [nenv][32][32] dense mass matrix (zero-initialised once; K0 writes the non-zeros)
  T* xkin;                        // [nenv][DEX_MAXBODY][16] body poses of K0: xpos(3) xquat(4) xmat(9)
  T* stage;                       // [nenv][DEX_MAXSLOT][8] contact staging slots of K1a / K1b (pair < 0 = no contact)
  int* nslot;                     // [nenv] staging slots in use
  int* work;                      // [work_cap][2] global narrowphase work list: env, slot << 24 | pair
  int* nwork;                     // [8] work items, fetch cursor of K1b, envs per solver-cost class (zeroed by K0 every mj_step)
  int* order;                     // optional [4][order_stride]: env lists per solver-cost class written by K1c (null = identity order)
  int order_stride;
  int work_cap;
  T* tree;                        // [nenv][8][32] K0 outputs: cdof rows 0-5, qfrc_smooth row 6, tree reference points row 7
  int cls_mode;                   // K1c's solver classes: 0 by contact count, 1 dense / structured (dex_arrow.cuh), 2 dense / sparse (dex_sparse.cuh)
  int skip_mode;                  // which envs a launch works on (see dex_env_off): 0 = those with skip == 0, 1 = those being placed (skip >= 2)
  double* time64;                 // optional [nenv][2]: data.time accumulated in double (MuJoCo's own arithmetic), solve-start time of the goal
};

// skip[env]: 0 = stepping normally, 1 = reset during this call (physics leaves it alone), 2 + k = reset and its prop pose is
// still being collision-checked, k rejected poses so far (the rounds of dex_place_env, dex_task.cuh).
template <typename T> DEX_DEV bool dex_env_off(const BatchState<T>& st, int env) {
  if (!st.skip) return false;
  const unsigned char s = st.skip[env];
  return st.skip_mode ? s < 2 : s != 0;
}

// ---------------------------------------------------------------------------------------
// math helpers
// ---------------------------------------------------------------------------------------
template <typename T> struct Real;
template <> struct Real<float> {
  static DEX_HD float eps() { return 1.1920929e-7f; }
  static DEX_HD float minval() { return 1e-15f; }
  static DEX_HD float big() { return 3.0e38f; }
};
template <> struct Real<double> {
  static DEX_HD double eps() { return 2.220446049250313e-16; }
  static DEX_HD double minval() { return 1e-15; }
  static DEX_HD double big() { return 1.0e300; }
};

template <typename T> DEX_DEV T dx_sqrt(T x) { return sqrt(x); }
// reciprocal square root: one MUFU.RSQ (+ one Newton step) in FP32 instead of IEEE sqrt + divide
#if defined(__CUDACC__)
DEX_DEV float dx_rsqrt(float x) {   // callers pass normal, positive x: the bare MUFU.RSQ (no denormal rescaling) + one Newton step
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * (1.5f - 0.5f * x * r * r);
}
DEX_DEV double dx_rsqrt(double x) { return 1.0 / sqrt(x); }
#else
template <typename T> inline T dx_rsqrt(T x) { return T(1) / std::sqrt(x); }
#endif
// reciprocal: MUFU.RCP + one Newton step in FP32 (positive, normal arguments), IEEE division in FP64
#if defined(__CUDACC__)
DEX_DEV float dx_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * (2.0f - x * r);
}
DEX_DEV double dx_rcp(double x) { return 1.0 / x; }
#else
template <typename T> inline T dx_rcp(T x) { return T(1) / x; }
#endif
template <typename T> DEX_DEV T dx_abs(T x) { return x < 0 ? -x : x; }
template <typename T> DEX_DEV T dx_min(T a, T b) { return a < b ? a : b; }
template <typename T> DEX_DEV T dx_max(T a, T b) { return a > b ? a : b; }
template <typename T> DEX_DEV T dx_clamp(T x, T lo, T hi) { return dx_min(dx_max(x, lo), hi); }

template <typename T> DEX_DEV T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename T> DEX_DEV void cross3(T* r, const T* a, const T* b) {
  T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> DEX_DEV T normalize3(T* a) {
  const T n2 = dot3(a, a);
  if (n2 < Real<T>::minval() * Real<T>::minval()) { a[0] = 1; a[1] = 0; a[2] = 0; return T(0); }
  const T inv = dx_rsqrt(n2);
  a[0] *= inv; a[1] *= inv; a[2] *= inv;
  return n2 * inv;
}
template <typename T> DEX_DEV void quat_mul(T* r, const T* a, const T* b) {
  T w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  T x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  T y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  T z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
template <typename T> DEX_DEV void quat_normalize(T* q) {
  const T n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (n2 < Real<T>::minval() * Real<T>::minval()) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
  else { const T inv = dx_rsqrt(n2); q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv; }
}
template <typename T> DEX_DEV void quat2mat(T* m, const T* q) {
  T w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
template <typename T> DEX_DEV void rot_vec(T* r, const T* m, const T* v) {
  T x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  T y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  T z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> DEX_DEV void rotT_vec(T* r, const T* m, const T* v) {
  T x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  T y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  T z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> DEX_DEV void axisangle2quat(T* q, const T* axis, T ang) {
  T s = sin(ang * T(0.5));
  q[0] = cos(ang * T(0.5)); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
// mju_makeFrame: frame[0:3] given (normal); builds the two tangents
template <typename T> DEX_DEV void make_frame(T* f) {
  normalize3(f);
  f[3] = 0; f[4] = 0; f[5] = 0;
  if (f[1] < T(0.5) && f[1] > T(-0.5)) f[4] = 1; else f[5] = 1;
  T dp = dot3(f, f + 3);
  f[3] -= dp * f[0]; f[4] -= dp * f[1]; f[5] -= dp * f[2];
  normalize3(f + 3);
  cross3(f + 6, f, f + 3);
}
// spatial algebra on [ang(3); lin(3)] vectors and 10-float inertias (MuJoCo layout)
template <typename T> DEX_DEV void mul_inert_vec(T* r, const T* c, const T* v) {
  r[0] = c[0] * v[0] + c[3] * v[1] + c[4] * v[2] - c[8] * v[4] + c[7] * v[5];
  r[1] = c[3] * v[0] + c[1] * v[1] + c[5] * v[2] + c[8] * v[3] - c[6] * v[5];
  r[2] = c[4] * v[0] + c[5] * v[1] + c[2] * v[2] - c[7] * v[3] + c[6] * v[4];
  r[3] = c[8] * v[1] - c[7] * v[2] + c[9] * v[3];
  r[4] = c[6] * v[2] - c[8] * v[0] + c[9] * v[4];
  r[5] = c[7] * v[0] - c[6] * v[1] + c[9] * v[5];
}
template <typename T> DEX_DEV void cross_motion(T* r, const T* v, const T* s) {
  T a[3], b[3], c[3];
  cross3(a, v, s);
  cross3(b, v, s + 3);
  cross3(c, v + 3, s);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
  r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
template <typename T> DEX_DEV void cross_force(T* r, const T* v, const T* f) {
  T a[3], b[3], c[3];
  cross3(a, v, f);
  cross3(b, v + 3, f + 3);
  cross3(c, v, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}

#endif  // DEX_MODEL_CUH_
