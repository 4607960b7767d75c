// dex_dynamics.cuh — K2: warp-per-env constraint assembly, Newton solve and semi-implicit Euler
// (restates SURVEY.md §8a rows 15 and 19-20): mj_makeConstraint + mj_makeImpedance +
// mj_referenceConstraint (dof friction loss, joint limits, pyramidal contacts), mj_solNewton (exact
// Hessian M + J^T D J, warp Cholesky, exact line search), mj_Euler with implicit joint damping.
// The smooth dynamics (cdof, M, qfrc_smooth) are produced by K0 (dex_tree.cuh).
//
// Lane mapping: lane = dof for every per-dof vector (kept in registers), lane = contact for
// per-contact scalars.  The contact-frame Jacobian Jc[c][3][nv] lives in a global (L2) scratch; rows
// of friction loss and joint limits are +-unit vectors and are never materialised.
#ifndef DEX_DYNAMICS_CUH_
#define DEX_DYNAMICS_CUH_

#include "dex_kin.cuh"
#include "dex_tree.cuh"

#define DEX_MSTRIDE 36   // rows 16-byte aligned; LDS.128 of 8 consecutive rows is bank-conflict free
// shared memory per warp (units of T): small per-dof vectors plus the Cholesky tile and the per-contact
// records.  The tree quantities (cdof, qfrc_smooth, M) come from K0 (dex_tree.cuh) through global
// scratch; the contact-frame Jacobian lives in an L2-resident global scratch (BatchState::jc).
#define DEX_K2_OFF_REF 0
#define DEX_K2_OFF_QPOS (DEX_K2_OFF_REF + 16)
#define DEX_K2_OFF_QVEL (DEX_K2_OFF_QPOS + 48)
#define DEX_K2_OFF_VECA (DEX_K2_OFF_QVEL + 32)
#define DEX_K2_OFF_VECB (DEX_K2_OFF_VECA + 32)
#define DEX_K2_OFF_H (DEX_K2_OFF_VECB + 32)
#define DEX_K2_OFF_CDAT (DEX_K2_OFF_H + 32 * DEX_MSTRIDE)
#define DEX_CDAT_STRIDE 24
#define DEX_K2_OFF_CDOT (DEX_K2_OFF_CDAT + DEX_MAXCON * DEX_CDAT_STRIDE)
#ifdef DEX_K2_M_IN_SMEM
#define DEX_K2_OFF_M (DEX_K2_OFF_CDOT + 96)
#define DEX_K2_SMEM_T (DEX_K2_OFF_M + 32 * 32)
#else
#define DEX_K2_SMEM_T (DEX_K2_OFF_CDOT + 96)
#endif
#define DEX_JC_ROW 32                         // global Jc scratch: [env][3 * DEX_MAXCON][DEX_JC_ROW]
#define DEX_JC_ENV (3 * DEX_MAXCON * DEX_JC_ROW)

// Timing experiments (tools/k2_breakdown.sh builds variants with these; never defined in the product build): DEX_EXP_FIXITER = a
// fixed number of Newton iterations with a refactorisation in each, DEX_EXP_NO{EULER,FACTOR,APPLY,MATVEC,LS} drop one part of
// the solver, and integrate == 2 (DEX_EXP, DEX_FREEZE=1 at run time) computes the step but keeps the state, so that every variant
// sees the same workload whatever it computes.
#ifdef DEX_EXP_FIXITER
#define DEX_EXP_ON true
#else
#define DEX_EXP_ON false
#endif

// cdat layout per contact
enum { CD_MU = 0, CD_D = 1, CD_AREF = 2 /*4*/, CD_JAR = 6 /*4*/, CD_JV = 10 /*4*/, CD_F = 14 /*3 frame-space force*/,
       CD_W = 17 /*5: W00 W01 W02 W11 W22*/, CD_MASK = 22 /*dof mask as int bits*/ };

// optional per-env debug outputs (null = off); used by the parity tests
template <typename T>
struct DynDebug {
  T* M;            // [nenv][nv*nv]
  T* qfrc_smooth;  // [nenv][nv]
  T* qacc;         // [nenv][nv]  solver output
  T* qfrc_constraint;  // [nenv][nv]
  T* efc;          // [nenv][DEX_MAXCON][8]: D, aref[4], mu, imp, R
};

template <typename T>
DEX_DEV_NOINLINE void dex_impedance(const T* solimp, T pos, T margin, T* imp) {  // getimpedance
  T dmin = dx_clamp(solimp[0], T(0.0001), T(0.9999));
  T dmax = dx_clamp(solimp[1], T(0.0001), T(0.9999));
  T width = dx_max(solimp[2], Real<T>::minval());
  T mid = dx_clamp(solimp[3], T(0.0001), T(0.9999));
  T power = dx_max(solimp[4], T(1));
  T x = dx_abs((pos - margin) / width);
  if (x >= 1) { *imp = dmax; return; }
  if (x <= 0) { *imp = dmin; return; }
  T y;
  if (power == 1) y = x;
  else if (power == 2) {   // MuJoCo's default solimp: no pow()
    const T a = x <= mid ? x / mid : (1 - x) / (1 - mid);
    y = x <= mid ? a * a * mid : 1 - a * a * (1 - mid);
  }
  else if (x <= mid) y = pow(x / mid, power) * mid;
  else y = 1 - pow((1 - x) / (1 - mid), power) * (1 - mid);
  *imp = dmin + y * (dmax - dmin);
}
template <typename T>
DEX_DEV void dex_kb(const T* solref, const T* solimp, T timestep, T* K, T* B) {
  T dmax = dx_clamp(solimp[1], T(0.0001), T(0.9999));
  T tc = solref[0], dr = solref[1];
  if (tc > 0) {
    tc = dx_max(tc, 2 * timestep);  // refsafe
    *K = 1 / dx_max(Real<T>::minval(), dmax * dmax * tc * tc * dr * dr);
    *B = 2 / dx_max(Real<T>::minval(), dmax * tc);
  } else {
    *K = -tc / dx_max(Real<T>::minval(), dmax * dmax);
    *B = -dr / dx_max(Real<T>::minval(), dmax);
  }
}

template <typename T> struct alignas(16) DexVec4 { T x, y, z, w; };
// Cholesky of an SPD matrix of order n <= 32, lane = row.  Storage contract (shared by the factor and
// the solves): the warp's 32 x DEX_MSTRIDE tile H holds the STRICTLY LOWER triangle only; the
// diagonal, the upper triangle and every row / column >= n are zero; lane i carries its diagonal
// entry in the register `hd` (1 for lanes >= n).  Because the upper part is zero, every dot product
// and every substitution step runs over whole 4-column groups without masks.
// Left-looking over 4-column blocks: the prefix dot products of the block's 4 columns share the
// loads of row i (5 LDS.128 per 16 FMAs), the 4 x 4 diagonal block is finished with shuffles, and
// one STS.128 per lane writes the block back; 8 warp barriers per factorisation (r01f profile: the
// column-at-a-time version cost 3.9 k warp instructions per call, 42 % of K2).
// `split` (a multiple of 4, or >= n): rows / columns >= split are known to be decoupled from those below it
// (block-diagonal matrix), so their prefix dot products start at column `split`.
// Returns 1 / L[i][i] for lane i.
template <typename T>
DEX_DEV_NOINLINE T dex_chol_factor(T* H, T hd, int n, int split, int* fail) {
  const int i = dex_lane();
  T* rowi = H + i * DEX_MSTRIDE;
  bool bad = false;
  T inv_mine = 1;
  const int nblk = (n + 3) >> 2;
  for (int kb = 0; kb < nblk; ++kb) {
    const int k0 = 4 * kb;
    const DexVec4<T> a = *reinterpret_cast<const DexVec4<T>*>(rowi + k0);
    T s0 = i == k0 ? hd : a.x, s1 = i == k0 + 1 ? hd : a.y, s2 = i == k0 + 2 ? hd : a.z, s3 = i == k0 + 3 ? hd : a.w;
    const T* r0 = H + k0 * DEX_MSTRIDE;
    // NOT unrolled (like the substitution loops below): short run-time trip counts, and the solver is sensitive to its code
    // footprint ("Code size": unroll 1 / 2 / 4 = 0.367 / 0.381 / 0.383 ms per launch)
#pragma unroll 1
    for (int g = (k0 >= split ? split : 0); g < k0; g += 4) {
      const DexVec4<T> vi = *reinterpret_cast<const DexVec4<T>*>(rowi + g);
      const DexVec4<T> v0 = *reinterpret_cast<const DexVec4<T>*>(r0 + g);
      const DexVec4<T> v1 = *reinterpret_cast<const DexVec4<T>*>(r0 + DEX_MSTRIDE + g);
      const DexVec4<T> v2 = *reinterpret_cast<const DexVec4<T>*>(r0 + 2 * DEX_MSTRIDE + g);
      const DexVec4<T> v3 = *reinterpret_cast<const DexVec4<T>*>(r0 + 3 * DEX_MSTRIDE + g);
      s0 -= vi.x * v0.x + vi.y * v0.y + vi.z * v0.z + vi.w * v0.w;
      s1 -= vi.x * v1.x + vi.y * v1.y + vi.z * v1.z + vi.w * v1.w;
      s2 -= vi.x * v2.x + vi.y * v2.y + vi.z * v2.z + vi.w * v2.w;
      s3 -= vi.x * v3.x + vi.y * v3.y + vi.z * v3.z + vi.w * v3.w;
    }
    // 4 x 4 diagonal block: column k0 + r is final once the r earlier columns of the block are applied
    // The pivots' square roots are kept OFF the critical path: column k0 + r is updated with s_k S_k / d_k (the pre-pivot
    // column values and a reciprocal) instead of l_k L_k, so a pivot's chain is shuffle -> reciprocal -> fused multiply-add;
    // the four rsqrt's and the products l_k = s_k / sqrt(d_k) follow independently at the end (same number of shuffles;
    // env step 3.94 -> 3.92 ms).
    T d0 = dex_shfl(s0, k0);
    const T a01 = dex_shfl(s0, k0 + 1), a02 = dex_shfl(s0, k0 + 2), a03 = dex_shfl(s0, k0 + 3);
    bad |= !(d0 > Real<T>::minval()); d0 = dx_max(d0, Real<T>::minval());
    const T q0 = s0 * dx_rcp(d0);
    s1 -= q0 * a01; s2 -= q0 * a02; s3 -= q0 * a03;
    T d1 = dex_shfl(s1, k0 + 1);
    const T a12 = dex_shfl(s1, k0 + 2), a13 = dex_shfl(s1, k0 + 3);
    bad |= !(d1 > Real<T>::minval()); d1 = dx_max(d1, Real<T>::minval());
    const T q1 = s1 * dx_rcp(d1);
    s2 -= q1 * a12; s3 -= q1 * a13;
    T d2 = dex_shfl(s2, k0 + 2);
    const T a23 = dex_shfl(s2, k0 + 3);
    bad |= !(d2 > Real<T>::minval()); d2 = dx_max(d2, Real<T>::minval());
    s3 -= s2 * dx_rcp(d2) * a23;
    T d3 = dex_shfl(s3, k0 + 3);
    bad |= !(d3 > Real<T>::minval()); d3 = dx_max(d3, Real<T>::minval());
    const T i0 = dx_rsqrt(d0), i1 = dx_rsqrt(d1), i2 = dx_rsqrt(d2), i3 = dx_rsqrt(d3);
    const T l0 = s0 * i0, l1 = s1 * i1, l2 = s2 * i2, l3 = s3 * i3;
    DexVec4<T> o;
    o.x = i > k0 ? l0 : T(0); o.y = i > k0 + 1 ? l1 : T(0); o.z = i > k0 + 2 ? l2 : T(0); o.w = i > k0 + 3 ? l3 : T(0);
    *reinterpret_cast<DexVec4<T>*>(rowi + k0) = o;
    inv_mine = i == k0 ? i0 : inv_mine;       // lane i keeps the inverse pivot of its own column (four selects)
    inv_mine = i == k0 + 1 ? i1 : inv_mine;
    inv_mine = i == k0 + 2 ? i2 : inv_mine;
    inv_mine = i == k0 + 3 ? i3 : inv_mine;
    dex_syncwarp();
  }
  if (bad) *fail = 1;
  return inv_mine;
}
// Solves L L^T x = b with the factor left in H by dex_chol_factor; lane i holds b_i / returns x_i.
// Both substitutions are unmasked thanks to the zero diagonal / upper triangle of the tile.
template <typename T>
DEX_DEV_NOINLINE T dex_chol_apply(const T* H, int n, T inv_mine, T b) {
  const int i = dex_lane();
  const T* rowi = H + i * DEX_MSTRIDE;
  const int n4 = (n + 3) & ~3;
#pragma unroll 1
  for (int k = 0; k < n4; k += 4) {   // L y = b
    const DexVec4<T> l = *reinterpret_cast<const DexVec4<T>*>(rowi + k);
    b -= l.x * dex_shfl(b * inv_mine, k);
    b -= l.y * dex_shfl(b * inv_mine, k + 1);
    b -= l.z * dex_shfl(b * inv_mine, k + 2);
    b -= l.w * dex_shfl(b * inv_mine, k + 3);
  }
  b *= inv_mine;
  const T* col = H + i;
#pragma unroll 1
  for (int k = n4 - 1; k >= 0; k -= 4) {   // L^T x = y
    const T c0 = col[k * DEX_MSTRIDE], c1 = col[(k - 1) * DEX_MSTRIDE], c2 = col[(k - 2) * DEX_MSTRIDE], c3 = col[(k - 3) * DEX_MSTRIDE];
    b -= c0 * dex_shfl(b * inv_mine, k);
    b -= c1 * dex_shfl(b * inv_mine, k - 1);
    b -= c2 * dex_shfl(b * inv_mine, k - 2);
    b -= c3 * dex_shfl(b * inv_mine, k - 3);
  }
  return b * inv_mine;
}
// Fills the warp's Cholesky tile from a symmetric matrix stored with rows of 32 (global scratch):
// strictly lower part of the leading n x n block, zeros elsewhere; returns this lane's diagonal entry.
template <typename T>
DEX_DEV_NOINLINE T dex_chol_load(T* H, const T* Mg, int n) {
  const int lane = dex_lane();
  // four rows per trip: lane = (row within the group, 4-column group); one 16-byte load and one 16-byte store per lane
  const int r = lane >> 3, c4 = (lane & 7) << 2;
#pragma unroll
  for (int j0 = 0; j0 < 32; j0 += 4) {
    const int j = j0 + r;
    DexVec4<T> v = {T(0), T(0), T(0), T(0)};
    if (j < n) v = *reinterpret_cast<const DexVec4<T>*>(Mg + j * 32 + c4);
    v.x = c4 < j ? v.x : T(0); v.y = c4 + 1 < j ? v.y : T(0); v.z = c4 + 2 < j ? v.z : T(0); v.w = c4 + 3 < j ? v.w : T(0);
    *reinterpret_cast<DexVec4<T>*>(H + j * DEX_MSTRIDE + c4) = v;
  }
  return lane < n ? Mg[lane * 32 + lane] : T(1);
}
// y_i = sum_j M[j][i] x_j (M symmetric) with M in the env's global scratch (rows of 32, read coalesced
// with lane = column) and x in shared memory.
template <typename T>
DEX_DEV_NOINLINE T dex_matvec(const T* Mg, const T* x, int nv) {
  const int i = dex_lane();
  T s0 = 0, s1 = 0;
  // compile-time trip count (predicated on j < nv) so that the loop is always fully unrolled and all row loads of the
  // product are in flight at once (L2 latency; partial unrolls of the run-time loop: 1 / 4 / 8 / full = 0.392 / 0.367 / 0.373 /
  // 0.357 ms per launch)
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const T a0 = j < nv ? Mg[j * 32 + i] : T(0), a1 = j + 1 < nv ? Mg[(j + 1) * 32 + i] : T(0);
    s0 += a0 * x[j]; s1 += a1 * x[j + 1];
  }
  return i < nv ? s0 + s1 : T(0);
}
// out[t] = Jc[t] . x for the 3*ncon contact-frame Jacobian rows.  Lane = ROW (rows t, t + 32, t + 64): every lane walks its own
// row of the env's global (L2 / L1) scratch with 16-byte loads against x in shared memory (32 entries, zero beyond nv).
// 48 instructions per 32 rows; the earlier lane = dof form (three rows per round, one warp reduction each) cost 40 per
// CONTACT, which made the envs with 10-20 contacts the stragglers of the whole launch (r02 niter probe).
template <typename T>
DEX_DEV_NOINLINE void dex_contact_dots(const T* Jc, const T* xs, int nrows, T* out) {
  const int lane = dex_lane();
#pragma unroll 1
  for (int t0 = 0; t0 < nrows; t0 += 32) {
    const int t = t0 + lane;
    const T* row = Jc + (t < nrows ? t : 0) * DEX_JC_ROW;
    T s = 0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      T a[4], x[4];
      dex_ld4(row + j, a); dex_ld4(xs + j, x);
      s += a[0] * x[0] + a[1] * x[1] + a[2] * x[2] + a[3] * x[3];
    }
    if (t < nrows) out[t] = s;
  }
  dex_syncwarp();
}

// H += Jc^T W Jc for one contact, strictly lower triangle of the warp's tile (the diagonal lives in registers, the caller
// adds it).  The contact touches n = popc(mask) dofs, i.e. n (n - 1) / 2 entries below the diagonal; they are dealt out 32
// to a round, lane p taking the pair (r, s), r > s, of the mask's dof list and fetching t = W a of dof r and a of dof s with
// six shuffles.  With the usual 13 dofs (finger chain, wrist, object) that is three rounds instead of the 13 trips, one per
// dof, of the earlier "every row walks the whole mask" loop.  `scratch`: 32 ints of shared memory (the dof list).
template <typename T>
DEX_DEV void dex_contact_hessian(T* H, unsigned mask, T a0, T a1, T a2, T t0, T t1, T t2, int* scratch) {
  const int lane = dex_lane();
  const int n = dex_popc(mask);
  if ((mask >> lane) & 1u) scratch[dex_popc(mask & ((1u << lane) - 1u))] = lane;
  dex_syncwarp();
  const int npair = n * (n - 1) / 2;
#pragma unroll 1
  for (int p0 = 0; p0 < npair; p0 += 32) {
    const int p = p0 + lane;
    const bool on = p < npair;
    // p = r (r - 1) / 2 + s, 0 <= s < r
    int r = (int)((T(1) + dx_sqrt(T(8 * (on ? p : 0) + 1))) * T(0.5));
    if (r * (r - 1) / 2 > p) --r;
    if ((r + 1) * r / 2 <= p) ++r;
    const int sidx = (on ? p : 0) - r * (r - 1) / 2;
    const int i = on ? scratch[r] : 0, j = on ? scratch[sidx] : 0;
    const T ti0 = dex_shfl(t0, i), ti1 = dex_shfl(t1, i), ti2 = dex_shfl(t2, i);
    const T bj0 = dex_shfl(a0, j), bj1 = dex_shfl(a1, j), bj2 = dex_shfl(a2, j);
    if (on) H[i * DEX_MSTRIDE + j] += ti0 * bj0 + ti1 * bj1 + ti2 * bj2;
  }
  dex_syncwarp();
}

template <typename T>
DEX_DEV void dex_dynamics_env(const DevModel<T>& m, const BatchState<T>& st, const DynDebug<T>& dbg, int env,
                              T* smem, int integrate) {
  if (env >= st.nenv || dex_env_off(st, env)) { return; }  // warp-uniform
  const int lane = dex_lane();
  const int nv = m.nv;
  T* ref = smem + DEX_K2_OFF_REF;
  T* qpos = smem + DEX_K2_OFF_QPOS;
  T* qvel = smem + DEX_K2_OFF_QVEL;
  T* vecA = smem + DEX_K2_OFF_VECA;
  T* vecB = smem + DEX_K2_OFF_VECB;
#ifdef DEX_K2_M_IN_SMEM
  T* M = smem + DEX_K2_OFF_M;                 // dense symmetric M staged from K0's global scratch
  {
    const T* Mg = st.mscr + (size_t)env * DEX_MG_ENV;
    for (int j = 0; j < nv; ++j) M[j * 32 + lane] = Mg[j * 32 + lane];
  }
#else
  T* M = st.mscr + (size_t)env * DEX_MG_ENV;  // global scratch (L2 resident), symmetric, written by K0
  dex_prefetch_l1(M + lane * 32);             // row `lane` (128 B): the first M v product comes ~1 k instructions later
#endif
  T* H = smem + DEX_K2_OFF_H;
  T* Jc = st.jc + (size_t)env * DEX_JC_ENV;  // global scratch (L2 resident)
  T* cdat = smem + DEX_K2_OFF_CDAT;
  T* cdot = smem + DEX_K2_OFF_CDOT;
  const T h = m.timestep;
  unsigned status = 0;

  // ------------------------------------------------------------------ load state and the tree quantities of K0
  for (int i = lane; i < m.nq; i += 32) qpos[i] = st.qpos[(size_t)env * st.qpos_stride + i];
  const bool isdof = lane < nv;
  const T qv = isdof ? st.qvel[(size_t)env * st.qvel_stride + lane] : T(0);
  T qacc = isdof ? st.warm[(size_t)env * st.warm_stride + lane] : T(0);
  qvel[lane] = qv;
  const T* tree = st.tree + (size_t)env * DEX_TREE_ENV;
  T cd[6];
  for (int k = 0; k < 6; ++k) cd[k] = isdof ? tree[k * 32 + lane] : T(0);
  const T fs = isdof ? tree[DEX_TREE_ROW_FS * 32 + lane] : T(0);  // qfrc_smooth
  if (lane < 16) ref[lane] = tree[DEX_TREE_ROW_REF * 32 + lane];
  const T damping = isdof ? dex_ldg(m.dof_damping + lane) : T(0);
  dex_syncwarp();

  // ------------------------------------------------------------------ constraint rows
  // friction loss (lane = dof)
  // (state independent up to aref = -B qvel: R f, D = 1 / R and B come precomputed from the host, dof_flrow)
  T f_loss = 0, f_lim = Real<T>::big(), f_D = 0, f_aref = 0;
  if (isdof) {
    f_loss = dex_ldg(m.dof_frictionloss + lane);
    T fr[4];
    dex_ldg4(m.dof_flrow + 4 * lane, fr);
    if (f_loss > 0) { f_lim = fr[0]; f_D = fr[1]; f_aref = -fr[2] * qv; }
  }
  // joint limit (lane = dof of a limited hinge/slide joint; at most one side active)
  T l_sign = 0, l_D = 0, l_aref = 0;
  if (isdof) {
    const int j = dex_ldg(m.dof_jnt + lane);
    if (dex_ldg(m.jnt_limited + j) && dex_ldg(m.jnt_type + j) >= DEX_JNT_SLIDE) {
      const T value = qpos[dex_ldg(m.jnt_qposadr + j)];
      const T margin = dex_ldg(m.jnt_margin + j);
      const T dlo = value - dex_ldg(m.jnt_range + 2 * j), dhi = dex_ldg(m.jnt_range + 2 * j + 1) - value;
      T pos = 0;
      if (dlo < margin) { l_sign = 1; pos = dlo; }
      else if (dhi < margin) { l_sign = -1; pos = dhi; }
      if (l_sign != 0) {
        T si[5], sr[2], imp, K, B;
        for (int k = 0; k < 5; ++k) si[k] = dex_ldg(m.jnt_solimp + 5 * j + k);
        sr[0] = dex_ldg(m.jnt_solref + 2 * j); sr[1] = dex_ldg(m.jnt_solref + 2 * j + 1);
        dex_impedance(si, pos, margin, &imp);
        dex_kb(sr, si, h, &K, &B);
        T R = dx_max(Real<T>::minval(), (1 - imp) / imp * dex_ldg(m.dof_invweight0 + lane));
        l_D = 1 / R;
        l_aref = -B * (l_sign * qv) - K * imp * (pos - margin);
      }
    }
  }
  // contacts: frame Jacobians (lane = dof)
  int ncon = st.ncon[env];
  if (ncon > DEX_MAXCON) ncon = DEX_MAXCON;
  const T* cin = st.contact + (size_t)env * DEX_MAXCON * DEX_CON_STRIDE;
  for (int c = 0; c < ncon; ++c) {
    const T* cc = cin + c * DEX_CON_STRIDE;
    T frame[9] = {cc[4], cc[5], cc[6], 0, 0, 0, 0, 0, 0};
    make_frame(frame);
    const int pr = (int)cc[7];
    const int b1 = dex_ldg(m.geom_body + dex_ldg(m.pair_geom1 + pr)), b2 = dex_ldg(m.geom_body + dex_ldg(m.pair_geom2 + pr));
    const unsigned m1 = (unsigned)dex_ldg(m.body_dofmask + b1), m2 = (unsigned)dex_ldg(m.body_dofmask + b2);
    T j0 = 0, j1 = 0, j2 = 0;
    if (isdof) {
      const T sgn = T((m2 >> lane) & 1u) - T((m1 >> lane) & 1u);
      if (sgn != 0) {
        const T* rp = ref + 3 * dex_ldg(m.dof_tree + lane);
        T off[3] = {cc[1] - rp[0], cc[2] - rp[1], cc[3] - rp[2]}, v[3];
        cross3(v, cd, off);
        v[0] += cd[3]; v[1] += cd[4]; v[2] += cd[5];
        j0 = sgn * dot3(frame, v); j1 = sgn * dot3(frame + 3, v); j2 = sgn * dot3(frame + 6, v);
      }
    }
    Jc[(3 * c) * DEX_JC_ROW + lane] = j0;
    Jc[(3 * c + 1) * DEX_JC_ROW + lane] = j1;
    Jc[(3 * c + 2) * DEX_JC_ROW + lane] = j2;
    if (lane == 0) *(int*)(cdat + c * DEX_CDAT_STRIDE + CD_MASK) = (int)(m1 | m2);
  }
  dex_syncwarp();
  dex_contact_dots(Jc, qvel, 3 * ncon, cdot);  // contact-frame velocities (qvel[] is zero beyond nv)
  if (lane < ncon) {  // lane = contact: impedance, R, aref of the 4 pyramid rows
    const int c = lane;
    const T* cc = cin + c * DEX_CON_STRIDE;
    const int pr = (int)cc[7];
    const int b1 = dex_ldg(m.geom_body + dex_ldg(m.pair_geom1 + pr)), b2 = dex_ldg(m.geom_body + dex_ldg(m.pair_geom2 + pr));
    const T mu_pair = dex_ldg(m.pair_friction + 5 * pr);
    const T cmargin = dex_ldg(m.pair_margin + pr) - dex_ldg(m.pair_gap + pr);
    T si[5], sr[2], imp, K, B;
    for (int k = 0; k < 5; ++k) si[k] = dex_ldg(m.pair_solimp + 5 * pr + k);
    sr[0] = dex_ldg(m.pair_solref + 2 * pr); sr[1] = dex_ldg(m.pair_solref + 2 * pr + 1);
    dex_impedance(si, cc[0], cmargin, &imp);
    dex_kb(sr, si, h, &K, &B);
    const T tran = dex_ldg(m.body_invweight0 + 2 * b1) + dex_ldg(m.body_invweight0 + 2 * b2);
    // condim 1: one frictionless row, R from diagApprox = tran, carried as mu = 0 (the per-contact loops below then take one
    // row instead of the four pyramid edges); condim 3: pyramid edges n +- mu t1, n +- mu t2 sharing Rpy = 2 mu^2 R0
    const bool frictionless = dex_ldg(m.pair_condim + pr) == 1;
    const T mu = frictionless ? T(0) : mu_pair;
    const T R0 = dx_max(Real<T>::minval(), (1 - imp) / imp * (tran + mu * mu * tran));
    const T Rpy = frictionless ? R0 : 2 * mu * mu * R0;
    T* cdt = cdat + c * DEX_CDAT_STRIDE;
    cdt[CD_MU] = mu;
    cdt[CD_D] = 1 / Rpy;
    const T vn = cdot[3 * c], v1 = cdot[3 * c + 1], v2 = cdot[3 * c + 2];
    const T kp = K * imp * (cc[0] - cmargin);
    cdt[CD_AREF + 0] = -B * (vn + mu * v1) - kp;
    cdt[CD_AREF + 1] = -B * (vn - mu * v1) - kp;
    cdt[CD_AREF + 2] = -B * (vn + mu * v2) - kp;
    cdt[CD_AREF + 3] = -B * (vn - mu * v2) - kp;
    if (dbg.efc) {
      T* e = dbg.efc + ((size_t)env * DEX_MAXCON + c) * 8;
      e[0] = cdt[CD_D]; e[1] = cdt[CD_AREF]; e[2] = cdt[CD_AREF + 1]; e[3] = cdt[CD_AREF + 2]; e[4] = cdt[CD_AREF + 3];
      e[5] = mu; e[6] = imp; e[7] = Rpy;
    }
  }
  dex_syncwarp();

  // ------------------------------------------------------------------ Newton solver (mj_solNewton)
  const T scale = 1 / (m.meaninertia * (nv > 1 ? nv : 1));
  const T tol = dx_max(m.tolerance, sizeof(T) == 4 ? T(1e-6) : T(1e-12));
  vecA[lane] = qacc;
  dex_syncwarp();
  T Ma = dex_matvec(M, vecA, nv);
  dex_contact_dots(Jc, vecA, 3 * ncon, cdot);
  if (lane < ncon) {
    T* cdt = cdat + lane * DEX_CDAT_STRIDE;
    const T mu = cdt[CD_MU], an = cdot[3 * lane], a1 = cdot[3 * lane + 1], a2 = cdot[3 * lane + 2];
    cdt[CD_JAR + 0] = an + mu * a1 - cdt[CD_AREF + 0];
    cdt[CD_JAR + 1] = an - mu * a1 - cdt[CD_AREF + 1];
    cdt[CD_JAR + 2] = an + mu * a2 - cdt[CD_AREF + 2];
    cdt[CD_JAR + 3] = an - mu * a2 - cdt[CD_AREF + 3];
  }
  T grad = 0, cost = 0, oldcost = 0, gnorm_prev = 0, hinv = 1, hdiag_prev = 0;
  int asig_prev = -1;
  bool small_step = false;
  int iter = 0;
  const int maxiter = m.iterations;
  for (;; ++iter) {
    // ---- forces, active sets, cost (PrimalUpdateConstraint)
    // per-dof rows without divergent branches (lanes >= nv carry qacc = f_loss = l_sign = 0): friction loss is
    // linear below -R f / above +R f and quadratic in between; a limit row is active while x < 0
    T mycost = T(0.5) * qacc * Ma - qacc * fs, fdof = 0, hdiag = 0;
    int asig = 0;
    {
      const T x = qacc - f_aref, lim = f_lim;
      const bool on = f_loss > 0, lo = x <= -lim, hi = x >= lim;
      const T c = lo ? f_loss * (T(-0.5) * f_lim - x) : (hi ? f_loss * (T(-0.5) * f_lim + x) : T(0.5) * f_D * x * x);
      const T f = lo ? f_loss : (hi ? -f_loss : -(f_D * x));
      mycost += on ? c : T(0); fdof += on ? f : T(0); hdiag += (on && !lo && !hi) ? f_D : T(0);
    }
    {
      const T x = l_sign * qacc - l_aref;
      const bool on = (l_sign != 0) && (x < 0);
      mycost += on ? T(0.5) * l_D * x * x : T(0); fdof -= on ? l_sign * l_D * x : T(0); hdiag += on ? l_D : T(0);
    }
    if (lane < ncon) {
      T* cdt = cdat + lane * DEX_CDAT_STRIDE;
      const T D = cdt[CD_D], mu = cdt[CD_MU];
      T f[4], w[4];
      const int nrow = mu == 0 ? 1 : 4;
      for (int r = 0; r < 4; ++r) {
        const T x = cdt[CD_JAR + r];
        const bool on = x < 0 && r < nrow;   // pyramid edge active
        f[r] = on ? -D * x : T(0); w[r] = on ? D : T(0); mycost += on ? T(0.5) * D * x * x : T(0); asig |= on ? 1 << r : 0;
      }
      cdt[CD_F + 0] = f[0] + f[1] + f[2] + f[3];
      cdt[CD_F + 1] = mu * (f[0] - f[1]);
      cdt[CD_F + 2] = mu * (f[2] - f[3]);
      cdt[CD_W + 0] = w[0] + w[1] + w[2] + w[3];
      cdt[CD_W + 1] = mu * (w[0] - w[1]);
      cdt[CD_W + 2] = mu * (w[2] - w[3]);
      cdt[CD_W + 3] = mu * mu * (w[0] + w[1]);
      cdt[CD_W + 4] = mu * mu * (w[2] + w[3]);
    }
    dex_syncwarp();
    oldcost = cost;
    cost = dex_warp_sum(mycost);
    // ---- gradient = M a - qfrc_smooth - J^T f
    T jtf = fdof;
    if (isdof) {
      for (int c = 0; c < ncon; ++c) {
        const T* cdt = cdat + c * DEX_CDAT_STRIDE;
        jtf += Jc[(3 * c) * DEX_JC_ROW + lane] * cdt[CD_F] + Jc[(3 * c + 1) * DEX_JC_ROW + lane] * cdt[CD_F + 1] +
               Jc[(3 * c + 2) * DEX_JC_ROW + lane] * cdt[CD_F + 2];
      }
    }
    grad = isdof ? (Ma - fs - jtf) : T(0);
    const T gnorm = dx_sqrt(dex_warp_sum(grad * grad));
#ifdef DEX_EXP_FIXITER
    if (iter >= DEX_EXP_FIXITER) break;
#else
    if (scale * gnorm < tol || small_step) break;
#endif
    if (iter > 0 && !DEX_EXP_ON) {
      // MuJoCo stops on improvement < tolerance; in FP32 the cost is only good to ~1e-7 relative,
      // so the improvement test is honoured only once the gradient has stopped shrinking.
      const T improvement = scale * (oldcost - cost);
      if (improvement < tol && (sizeof(T) == 8 || gnorm > T(0.5) * gnorm_prev)) break;
    }
    gnorm_prev = gnorm;
    if (iter >= maxiter) { status |= DEX_STATUS_SOLVER_MAXITER; break; }

    // ---- Hessian H = M + J^T D_active J, Cholesky, search = -H^{-1} grad.  The factor is reused
    // when the active set (hence H) is the same as in the previous iteration.
    const unsigned sig_changed = dex_ballot(asig != asig_prev || hdiag != hdiag_prev);
    asig_prev = asig; hdiag_prev = hdiag;
#ifdef DEX_EXP_NOFACTOR
    if (false) {
#else
    if (iter == 0 || sig_changed || DEX_EXP_ON) {
#endif
      T hd = dex_chol_load(H, M, nv) + hdiag;
      dex_syncwarp();
      bool cross = false;
      const unsigned splitlo = m.chol_split >= 32 ? 0xffffffffu : ((1u << m.chol_split) - 1u);
      for (int c = 0; c < ncon; ++c) {
        const T* cdt = cdat + c * DEX_CDAT_STRIDE;
        const T w00 = cdt[CD_W], w01 = cdt[CD_W + 1], w02 = cdt[CD_W + 2], w11 = cdt[CD_W + 3], w22 = cdt[CD_W + 4];
        if (w00 == 0) continue;  // no active row (warp-uniform)
        unsigned mask = (unsigned)*(const int*)(cdat + c * DEX_CDAT_STRIDE + CD_MASK);
        if ((mask & splitlo) && (mask & ~splitlo)) cross = true;   // couples the trees on both sides of chol_split
        const T a0 = Jc[(3 * c) * DEX_JC_ROW + lane], a1 = Jc[(3 * c + 1) * DEX_JC_ROW + lane], a2 = Jc[(3 * c + 2) * DEX_JC_ROW + lane];
        const T t0 = a0 * w00 + a1 * w01 + a2 * w02, t1 = a0 * w01 + a1 * w11, t2 = a0 * w02 + a2 * w22;
        const bool mine = isdof && ((mask >> lane) & 1u);
        if (mine) hd += t0 * a0 + t1 * a1 + t2 * a2;
        dex_contact_hessian(H, mask & (nv >= 32 ? 0xffffffffu : ((1u << nv) - 1u)), a0, a1, a2, t0, t1, t2, reinterpret_cast<int*>(cdot));
      }
      dex_syncwarp();
      int cfail = 0;
      hinv = dex_chol_factor(H, hd, nv, cross ? nv : m.chol_split, &cfail);
      if (cfail) status |= DEX_STATUS_CHOL_FAIL;
    }
#ifdef DEX_EXP_NOAPPLY
    const T search = isdof ? T(-1e-3) * grad : T(0);
#else
    const T search = dex_chol_apply(H, nv, hinv, isdof ? -grad : T(0));
#endif

    // ---- exact line search along `search` (PrimalSearch)
    vecB[lane] = isdof ? search : T(0);
    dex_syncwarp();
#ifdef DEX_EXP_NOMATVEC
    const T Mv = search;
#else
    const T Mv = dex_matvec(M, vecB, nv);
#endif
    dex_contact_dots(Jc, vecB, 3 * ncon, cdot);
    if (lane < ncon) {
      T* cdt = cdat + lane * DEX_CDAT_STRIDE;
      const T mu = cdt[CD_MU], sn = cdot[3 * lane], s1 = cdot[3 * lane + 1], s2 = cdot[3 * lane + 2];
      cdt[CD_JV + 0] = sn + mu * s1; cdt[CD_JV + 1] = sn - mu * s1;
      cdt[CD_JV + 2] = sn + mu * s2; cdt[CD_JV + 3] = sn - mu * s2;
    }
    dex_syncwarp();
    const T q1 = isdof ? search * (Ma - fs) : T(0), q2 = isdof ? search * Mv : T(0);
    // The first Newton step of the 1-D problem needs no row pass: at alpha = 0 the derivative is search . grad and the
    // curvature search . H search = -search . grad (H search = -grad), so the step lands on alpha = 1 exactly.
    const T g0 = dex_warp_sum(isdof ? search * grad : T(0));
    T alpha = g0 < 0 ? T(1) : T(0), lo = 0, hi = -1;   // not a descent direction: alpha stays 0
#ifdef DEX_EXP_NOLS
    for (int ls = 1; ls < 1 && g0 < 0; ++ls) {
#else
    for (int ls = 1; ls < 24 && g0 < 0; ++ls) {
#endif
      T g = q1 + alpha * q2, hh = q2;
      {  // friction loss and limit rows, branch-free (same regimes as in the force update above)
        const T x = qacc + alpha * search - f_aref, lim = f_lim;
        const bool on = f_loss > 0, lo = x <= -lim, hi = x >= lim;
        const T dg = lo ? -(f_loss * search) : (hi ? f_loss * search : f_D * x * search);
        g += on ? dg : T(0); hh += (on && !lo && !hi) ? f_D * search * search : T(0);
        const T jv = l_sign * search, xl = l_sign * qacc - l_aref + alpha * jv;
        const bool lon = (l_sign != 0) && (xl < 0);
        g += lon ? l_D * xl * jv : T(0); hh += lon ? l_D * jv * jv : T(0);
      }
      if (lane < ncon) {
        const T* cdt = cdat + lane * DEX_CDAT_STRIDE;
        const T D = cdt[CD_D];
        const int nrow = cdt[CD_MU] == 0 ? 1 : 4;
        for (int r = 0; r < 4; ++r) {
          const T jv = cdt[CD_JV + r], x = cdt[CD_JAR + r] + alpha * jv;
          const bool on = x < 0 && r < nrow;
          g += on ? D * x * jv : T(0); hh += on ? D * jv * jv : T(0);
        }
      }
      g = dex_warp_sum(g);
      hh = dex_warp_sum(hh);
      if (dx_abs(g) <= (sizeof(T) == 4 ? T(1e-3) : T(1e-12)) * dx_abs(g0)) break;
      if (g < 0) lo = alpha; else hi = alpha;
      T next = alpha - g / hh;
      if (hi >= 0 && (next <= lo || next >= hi)) next = T(0.5) * (lo + hi);
      // converged in alpha: FP32 cannot push |g| below its rounding noise once g0 is small, and the outer Newton iteration
      // absorbs an error of this size in the step length
      const bool conv = dx_abs(next - alpha) <= (sizeof(T) == 4 ? T(1e-4) : T(1e-12)) * dx_abs(next);
      alpha = next;
      if (conv) break;
    }
    if (alpha == 0 && !DEX_EXP_ON) break;
    // MuJoCo's stopping rule "improvement * scale < tolerance", with the improvement taken from the
    // line search itself (-alpha g0 / 2 at an exact line minimum) instead of a difference of costs,
    // which FP32 cannot resolve.  The forces/gradient are still refreshed at the final iterate.
    small_step = scale * (T(-0.5) * alpha * g0) < m.tolerance;
    // ---- move
    if (isdof) { qacc += alpha * search; Ma += alpha * Mv; }
    if (lane < ncon) {
      T* cdt = cdat + lane * DEX_CDAT_STRIDE;
      for (int r = 0; r < 4; ++r) cdt[CD_JAR + r] += alpha * cdt[CD_JV + r];
    }
    dex_syncwarp();
  }

  // ------------------------------------------------------------------ outputs of the forward pass
  const T qfrc_constraint = isdof ? (Ma - grad - fs) : T(0);  // = J^T f at the final iterate
  if (dbg.qacc && isdof) dbg.qacc[(size_t)env * nv + lane] = qacc;
  if (dbg.qfrc_smooth && isdof) dbg.qfrc_smooth[(size_t)env * nv + lane] = fs;
  if (dbg.qfrc_constraint && isdof) dbg.qfrc_constraint[(size_t)env * nv + lane] = qfrc_constraint;
  if (dbg.M && isdof) for (int j = 0; j < nv; ++j) dbg.M[((size_t)env * nv + lane) * nv + j] = M[j * 32 + lane];
  const bool bad = isdof && !(dx_abs(qacc) < T(1e10));  // mj_checkAcc (also catches NaN)
  if (dex_ballot(bad)) status |= DEX_STATUS_BADQACC;
  if (lane == 0) {
    st.niter[env] = iter;
    if (status) st.status[env] |= status;
  }
  if (!integrate) return;

  // ------------------------------------------------------------------ mj_Euler (implicit joint damping)
  // only the leading m.nv_euler dofs belong to kinematic trees with joint damping; for the rest
  // (e.g. the free object) M + hB = M is block-decoupled and the solver's qacc is already the answer
  const int ne = m.nv_euler;
#ifdef DEX_EXP_NOEULER
  T qacc_e = qacc + damping * fs;
#else
  const T ehd = dex_chol_load(H, M, ne) + (lane < ne ? h * damping : T(0));
  dex_syncwarp();
  int efail = 0;
  const T einv = dex_chol_factor(H, ehd, ne, ne, &efail);
  T qacc_e = dex_chol_apply(H, ne, einv, lane < ne ? (fs + qfrc_constraint) : T(0));
#endif
  if (lane >= ne) qacc_e = qacc;
  const T qv_new = qv + h * qacc_e;
  dex_syncwarp();
  qvel[lane] = qv_new;
  dex_syncwarp();
#ifdef DEX_EXP
  if (integrate == 2) { if (isdof && qv_new == T(1.2345e33)) st.niter[env] = -1; return; }   // dry run: the state stays as it is
#endif
  if (isdof) {
    const int j = dex_ldg(m.dof_jnt + lane), da = dex_ldg(m.jnt_dofadr + j), qa = dex_ldg(m.jnt_qposadr + j);
    if (dex_ldg(m.jnt_type + j) == DEX_JNT_FREE) {
      if (lane == da) {
        T* qp = st.qpos + (size_t)env * st.qpos_stride + qa;
        for (int k = 0; k < 3; ++k) qp[k] = qpos[qa + k] + h * qvel[da + k];
        T w[3] = {qvel[da + 3], qvel[da + 4], qvel[da + 5]};
        T q[4] = {qpos[qa + 3], qpos[qa + 4], qpos[qa + 5], qpos[qa + 6]};
        T n = dx_sqrt(dot3(w, w));
        if (n >= Real<T>::minval()) {  // mju_quatIntegrate
          T ax[3] = {w[0] / n, w[1] / n, w[2] / n}, dq[4], r[4];
          axisangle2quat(dq, ax, h * n);
          quat_mul(r, q, dq);
          quat_normalize(r);
          q[0] = r[0]; q[1] = r[1]; q[2] = r[2]; q[3] = r[3];
        }
        for (int k = 0; k < 4; ++k) qp[3 + k] = q[k];
      }
    } else {
      st.qpos[(size_t)env * st.qpos_stride + qa] = qpos[qa] + h * qv_new;
    }
    st.qvel[(size_t)env * st.qvel_stride + lane] = qv_new;
    st.warm[(size_t)env * st.warm_stride + lane] = qacc;
  }
  if (lane == 0) {
    if (st.time64) { const double t = st.time64[2 * (size_t)env] + m.timestep_d; st.time64[2 * (size_t)env] = t; st.time[env] = (T)t; }
    else st.time[env] += h;
  }
}

#endif  // DEX_DYNAMICS_CUH_
