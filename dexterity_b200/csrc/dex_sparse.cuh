// dex_sparse.cuh — sparse form of the solver kernel K2 (lane = dof, one warp per env): level-scheduled L^T D L.
//
// The Newton Hessian H = M + diag + J^T W J and the implicit-Euler matrix M + h B inherit the pattern of the mass matrix:
// a dof couples only to its ancestors in an elimination forest (MuJoCo factorises M the same way, mj_factorI: dof n - 1
// down to 0, every dof updating the rows of its ancestors, no fill-in).  Dofs at the same depth of different subtrees are
// independent, so the factorisation runs level by level, deepest first: 7 steps for the Shadow hand (wrist 2 + finger 5),
// 13 with the object's six dofs chained above the wrist, instead of the 30 dependent pivots of the dense Cholesky
// (dex_dynamics.cuh), and the triangular solves likewise.  K2 is bound by the latency of exactly these chains
// (tools/k2_breakdown.py: factorisations, Euler and substitutions are half of its time).
//
// Lane i keeps row i in registers, indexed by SLOT = depth of the ancestor: e[d] = H[i][ancestor at depth d], d < depth_i,
// and the diagonal in `ed`; slots are compile-time indices after unrolling, so nothing is dynamically indexed.  A lane that
// is eliminated publishes its row in shared memory (32 rows of DEX_SP_ROW words); its ancestors pull it from there.
// TRUNK lanes (the dofs above the branching point: wrist, or object + wrist) are ancestors of several rows per step; they
// take no part in the steps of the rows below them and receive those rows' total in one deferred update instead
// (S -= V^T D^-1 V over the published rows, one matrix entry per lane), and the same in the forward substitution.
// Which lanes are eliminated when and who pulls from whom is host-built (dex_host_model.h; DexSpConst, DEX_SPL_*):
// structure 0 = the dof forest, structure 1 = the free object's dof chain made the common ancestor of the hand's tree,
// which keeps the pattern closed under contacts between the object and one finger / the palm (K1c's classes 4-7).
#ifndef DEX_SPARSE_CUH_
#define DEX_SPARSE_CUH_

#include "dex_dynamics.cuh"

#define DEX_SP_ROW 20     // words per published row: 16 slots, 1 / D, z of the forward substitution, padding (16-byte aligned rows)
#define DEX_SP_MROW 12    // words per row of the mass matrix kept in shared memory: 8 slots of structure 0, diagonal, padding
// shared memory per warp (units of T)
#define DEX_K2S_OFF_ROWS DEX_K2_OFF_H
#define DEX_K2S_OFF_MROWS (DEX_K2S_OFF_ROWS + 32 * DEX_SP_ROW)
#define DEX_K2S_OFF_CDAT (DEX_K2S_OFF_MROWS + 32 * DEX_SP_MROW)
#define DEX_K2S_OFF_CDOT (DEX_K2S_OFF_CDAT + DEX_MAXCON * DEX_CDAT_STRIDE)
#define DEX_K2S_SMEM_T (DEX_K2S_OFF_CDOT + 96)
#define CD_DMAX 23   // cdat word: deepest slot the contact's dofs own (int)

template <typename T>
struct SpLane {
  int depth, step;       // slot of the own diagonal (= number of ancestors), elimination step
  int mshift;            // slot of this structure at which the lane's structure-0 row starts (0, or the chain length above its tree)
  int mdepth;            // the lane's depth in structure 0
  unsigned anc[4];       // ancestor lane per slot, one byte each (own lane as padding)
  unsigned long long desc_lo, desc_hi;   // row to pull at step s, one byte each (0xff: none)
  unsigned schur;        // entries of the deferred trunk update
  T e[DEX_SP_SLOTS];     // row entries by slot
  T ed;                  // diagonal; D_i once the lane is eliminated
  T dinv;                // 1 / D_i
};

template <typename T> DEX_DEV int sp_anc(const SpLane<T>& L, int d) { return (int)((L.anc[d >> 2] >> (8 * (d & 3))) & 0xffu); }
template <typename T> DEX_DEV int sp_desc(const SpLane<T>& L, int s) {
  return (int)(((s < 8 ? L.desc_lo : L.desc_hi) >> (8 * (s & 7))) & 0xffull);
}

template <typename T>
DEX_DEV void sp_init(SpLane<T>& L, const DevModel<T>& m, int structure) {
  const int lane = dex_lane();
  const int* tab = m.sp_tab + structure * (DEX_SPL_WORDS * 32) + lane;
  L.depth = dex_ldg(tab + DEX_SPL_DEPTH * 32);
  L.step = dex_ldg(tab + DEX_SPL_STEP * 32);
  L.mdepth = dex_ldg(m.sp_tab + DEX_SPL_DEPTH * 32 + lane);
  L.mshift = L.depth - L.mdepth;
  unsigned d[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { L.anc[k] = (unsigned)dex_ldg(tab + (DEX_SPL_ANC + k) * 32); d[k] = (unsigned)dex_ldg(tab + (DEX_SPL_DESC + k) * 32); }
  L.desc_lo = (unsigned long long)d[0] | ((unsigned long long)d[1] << 32);
  L.desc_hi = (unsigned long long)d[2] | ((unsigned long long)d[3] << 32);
  L.schur = (unsigned)dex_ldg(tab + DEX_SPL_SCHUR * 32);
}

// the mass matrix of K0 (global scratch, symmetric, rows of 32) as structure-0 rows in shared memory: once per mj_step
template <typename T>
DEX_DEV void sp_stage_M(const DevModel<T>& m, const T* Mg, T* mrows, int nv) {
  const int lane = dex_lane();
  const int* tab = m.sp_tab + lane;
  const int depth = dex_ldg(tab + DEX_SPL_DEPTH * 32);
  const unsigned a0 = (unsigned)dex_ldg(tab + DEX_SPL_ANC * 32), a1 = (unsigned)dex_ldg(tab + (DEX_SPL_ANC + 1) * 32);
  const T* row = Mg + (lane < nv ? lane : 0) * 32;
  T v[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    const int a = (int)(((d < 4 ? a0 : a1) >> (8 * (d & 3))) & 0xffu);
    v[d] = (d < depth && lane < nv) ? row[a] : T(0);
  }
  const T dg = lane < nv ? row[lane] : T(1);
  T* o = mrows + lane * DEX_SP_MROW;
#pragma unroll
  for (int d = 0; d < 8; ++d) o[d] = v[d];
  o[8] = dg;
}

// row of M (+ diag_add on the diagonal) in the slots of the lane's current structure
template <typename T>
DEX_DEV void sp_load(SpLane<T>& L, const T* mrows, T diag_add) {
  const T* o = mrows + dex_lane() * DEX_SP_MROW;
#pragma unroll
  for (int d = 0; d < DEX_SP_SLOTS; ++d) {
    const int k = d - L.mshift;
    L.e[d] = (k >= 0 && k < L.mdepth) ? o[k] : T(0);   // chained structure: M does not couple the lane to the chain above its tree
  }
  L.ed = o[8] + diag_add;
  L.dinv = 1;
}

// H += J^T W J of one contact: a0..a2 = this lane's column of the contact-frame Jacobian (0 outside the contact's dofs),
// t = W a.  `dmax` = deepest slot any dof of the contact owns (warp-uniform; slots >= dmax cannot receive anything).
template <typename T>
DEX_DEV void sp_add_contact(SpLane<T>& L, T a0, T a1, T a2, T t0, T t1, T t2, int dmax) {
#pragma unroll
  for (int d = 0; d < DEX_SP_SLOTS; ++d) {
    if (d < dmax) {
      const int a = sp_anc(L, d);
      const T b0 = dex_shfl(a0, a), b1 = dex_shfl(a1, a), b2 = dex_shfl(a2, a);
      if (d < L.depth) L.e[d] += t0 * b0 + t1 * b1 + t2 * b2;
    }
  }
  L.ed += t0 * a0 + t1 * a1 + t2 * a2;
}

// In-place L^T D L.  Afterwards `rows` holds, for every lane, its row at the time of its elimination and 1 / D
// (rows[k][slot] / D_k = L[k][ancestor]); sp_solve reads them.  Sets *fail on a non-positive pivot.
template <typename T>
DEX_DEV void sp_factor(SpLane<T>& L, const DexSpConst& C, T* rows, int* fail) {
  const int lane = dex_lane();
  T* mine = rows + lane * DEX_SP_ROW;
  bool bad = false;
#pragma unroll 1
  for (int s = 0; s < C.nstep; ++s) {
    if (s == C.first_trunk_step && C.ndef > 0) {
      // deferred update of the trunk rows: entry (t, j) -= sum over the rows k below the trunk of row_k[t] row_k[j] / D_k.
      // One entry per lane (36 entries for a trunk of 8: lanes 0-3 take a second one), then the trunk lanes collect theirs.
      const int t1 = L.schur & 15, j1 = (L.schur >> 4) & 15, t2 = (L.schur >> 8) & 15, j2 = (L.schur >> 12) & 15;
      const bool on1 = t1 != 15, on2 = t2 != 15;
      T acc1 = 0, acc2 = 0;
      const bool two = C.nt > 7;   // warp-uniform
#pragma unroll 2
      for (int n = 0; n < C.ndef; ++n) {
        const T* rk = rows + (int)C.def[n] * DEX_SP_ROW;
        const T w = rk[16];
        const T p1 = rk[on1 ? t1 : 0] * w;
        acc1 += p1 * rk[on1 ? j1 : 0];
        if (two) { const T p2 = rk[on2 ? t2 : 0] * w; acc2 += p2 * rk[on2 ? j2 : 0]; }
      }
      const bool trunk = L.depth < C.nt && lane == (int)C.trunk[L.depth < 8 ? L.depth : 0];
      const int q0 = L.depth * (L.depth + 1) / 2;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < C.nt) {
          const int q = q0 + j;
          const T v1 = dex_shfl(acc1, q & 31);
          T v = v1;
          if (two) { const T v2 = dex_shfl(acc2, q & 31); v = q >= 32 ? v2 : v1; }
          if (trunk && j < L.depth) L.e[j] -= v;
          if (trunk && j == L.depth) L.ed -= v;
        }
      }
    }
    if (L.step == s) {
      bad |= !(L.ed > Real<T>::minval());
      L.ed = dx_max(L.ed, Real<T>::minval());
      L.dinv = dx_rcp(L.ed);
      dex_store16(mine, L.e);
      mine[16] = L.dinv;
    }
    dex_syncwarp();
    const int k = sp_desc(L, s);
    if (k != 0xff) {
      const T* rk = rows + k * DEX_SP_ROW;
      const T c = rk[L.depth];      // H[k][i]
      const T tmp = c * rk[16];
#pragma unroll
      for (int g = 0; g < DEX_SP_SLOTS; g += 4) {
        if (g < L.depth) {
          T v[4];
          dex_ld4(rk + g, v);
          L.e[g] -= tmp * v[0]; L.e[g + 1] -= tmp * v[1]; L.e[g + 2] -= tmp * v[2]; L.e[g + 3] -= tmp * v[3];   // slots >= depth are never read
        }
      }
      L.ed -= tmp * c;
    }
  }
  dex_syncwarp();
  if (bad) *fail = 1;
}

// x = H^{-1} b with the factor of sp_factor (mj_solveLD: L^-T, D^-1, L^-1)
template <typename T>
DEX_DEV T sp_solve(const SpLane<T>& L, const DexSpConst& C, T* rows, T b) {
  const int lane = dex_lane();
  // x <- L^-T x, deepest dofs first: a dof's value is final when its step comes (all its descendants went before)
#pragma unroll 1
  for (int s = 0; s < C.nstep; ++s) {
    if (s == C.first_trunk_step && C.ndef > 0) {
      // the trunk's share of the rows below it: lane = (trunk slot t, quarter of the rows); z of every row through shared memory
      rows[lane * DEX_SP_ROW + 17] = b * L.dinv;
      dex_syncwarp();
      const int t = lane & 7, part = lane >> 3;
      T p = 0;
      for (int n = part; n < C.ndef; n += 4) {
        const T* rk = rows + (int)C.def[n] * DEX_SP_ROW;
        p += rk[t] * rk[17];
      }
      p += dex_shfl_xor(p, 8);
      p += dex_shfl_xor(p, 16);
      const T mine = dex_shfl(p, L.depth & 7);
      if (L.depth < C.nt && lane == (int)C.trunk[L.depth < 8 ? L.depth : 0]) b -= mine;
    }
    const int k = sp_desc(L, s);
    const T zk = dex_shfl(b * L.dinv, k == 0xff ? lane : k);
    if (k != 0xff) b -= rows[k * DEX_SP_ROW + L.depth] * zk;
  }
  // x <- D^-1 x, then x <- L^-1 x by slot: the ancestor at slot d is final once the slots below d are in
  T x = b * L.dinv, acc = 0;
#pragma unroll
  for (int d = 0; d < DEX_SP_SLOTS; ++d) {
    if (d < C.nslot) {
      if (L.depth == d) x -= L.dinv * acc;
      const T xa = dex_shfl(x, sp_anc(L, d));
      if (d < L.depth) acc += L.e[d] * xa;
    }
  }
  if (L.depth >= C.nslot) x -= L.dinv * acc;   // the deepest lanes
  return x;
}

// K2, sparse form.  Same constraint rows, Newton iteration, line search and Euler step as dex_dynamics_env (see there for the
// MuJoCo routines each part follows); the factorisations and substitutions are the sparse ones above, and the Euler solve
// shares their call sites with the Newton iteration (the solver is sensitive to its instruction footprint).
template <typename T>
DEX_DEV void dex_dynamics_sp_env(const DevModel<T>& m, const BatchState<T>& st, const DynDebug<T>& dbg, int env,
                                 T* smem, int integrate) {
  if (env >= st.nenv || dex_env_off(st, env)) { return; }  // warp-uniform
  const int lane = dex_lane();
  const int nv = m.nv;
  T* ref = smem + DEX_K2_OFF_REF;
  T* qpos = smem + DEX_K2_OFF_QPOS;
  T* qvel = smem + DEX_K2_OFF_QVEL;
  T* vecA = smem + DEX_K2_OFF_VECA;
  T* vecB = smem + DEX_K2_OFF_VECB;
  T* rows = smem + DEX_K2S_OFF_ROWS;
  T* mrows = smem + DEX_K2S_OFF_MROWS;
  T* cdat = smem + DEX_K2S_OFF_CDAT;
  T* cdot = smem + DEX_K2S_OFF_CDOT;
  T* M = st.mscr + (size_t)env * DEX_MG_ENV;  // global scratch (L2 resident), symmetric, written by K0
  T* Jc = st.jc + (size_t)env * DEX_JC_ENV;   // global scratch (L2 resident)
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
  int ncon = st.ncon[env];
  if (ncon > DEX_MAXCON) ncon = DEX_MAXCON;
  sp_stage_M(m, M, mrows, nv);
  SpLane<T> L;
  int structure = ncon > 0 ? 1 : 0;   // contact-free envs: the dof forest itself (half as many steps)
  sp_init(L, m, structure);
  dex_syncwarp();

  // ------------------------------------------------------------------ constraint rows
  // friction loss (lane = dof): state independent up to aref = -B qvel, precomputed on the host (dof_flrow)
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
    const int dmax = dex_warp_max((((m1 ^ m2) >> lane) & 1u) ? L.depth : 0);
    if (lane == 0) { *(int*)(cdat + c * DEX_CDAT_STRIDE + CD_MASK) = (int)(m1 | m2); *(int*)(cdat + c * DEX_CDAT_STRIDE + CD_DMAX) = dmax; }
  }
  dex_syncwarp();
  dex_contact_dots(Jc, qvel, 3 * ncon, cdot);  // contact-frame velocities (qvel[] is zero beyond nv)
  if (lane < ncon) {  // lane = contact: impedance, R, aref of the 4 pyramid rows
    const int c = lane;
    const T* cc = cin + c * DEX_CON_STRIDE;
    const int pr = (int)cc[7];
    const int b1 = dex_ldg(m.geom_body + dex_ldg(m.pair_geom1 + pr)), b2 = dex_ldg(m.geom_body + dex_ldg(m.pair_geom2 + pr));
    const T mu = dex_ldg(m.pair_friction + 5 * pr);
    const T cmargin = dex_ldg(m.pair_margin + pr) - dex_ldg(m.pair_gap + pr);
    T si[5], sr[2], imp, K, B;
    for (int k = 0; k < 5; ++k) si[k] = dex_ldg(m.pair_solimp + 5 * pr + k);
    sr[0] = dex_ldg(m.pair_solref + 2 * pr); sr[1] = dex_ldg(m.pair_solref + 2 * pr + 1);
    dex_impedance(si, cc[0], cmargin, &imp);
    dex_kb(sr, si, h, &K, &B);
    const T tran = dex_ldg(m.body_invweight0 + 2 * b1) + dex_ldg(m.body_invweight0 + 2 * b2);
    const T R0 = dx_max(Real<T>::minval(), (1 - imp) / imp * (tran + mu * mu * tran));
    const T Rpy = 2 * mu * mu * R0;
    T* cdt = cdat + c * DEX_CDAT_STRIDE;
    cdt[CD_MU] = mu;
    cdt[CD_D] = 1 / Rpy;
    const T vn = cdot[3 * c], v1 = cdot[3 * c + 1], v2 = cdot[3 * c + 2];
    const T kp = K * imp * (cc[0] - cmargin);
    cdt[CD_AREF + 0] = -B * (vn + mu * v1) - kp;
    cdt[CD_AREF + 1] = -B * (vn - mu * v1) - kp;
    cdt[CD_AREF + 2] = -B * (vn + mu * v2) - kp;
    cdt[CD_AREF + 3] = -B * (vn - mu * v2) - kp;
  }
  dex_syncwarp();

  // ------------------------------------------------------------------ Newton solver (mj_solNewton), then the Euler solve
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
  T grad = 0, cost = 0, oldcost = 0, gnorm_prev = 0, hdiag_prev = 0;
  T qfrc_constraint = 0, qacc_e = 0;
  int asig_prev = -1;
  bool small_step = false, euler = false, zero_step = false;
  int iter = 0;
  const int maxiter = m.iterations;
  for (;; ++iter) {
    // ---- forces, active sets, cost (PrimalUpdateConstraint); see dex_dynamics_env
    T mycost = T(0.5) * qacc * Ma - qacc * fs, fdof = 0, hdiag = 0;
    int asig = 0;
    {
      const T x = qacc - f_aref;
      const bool on = f_loss > 0, lo = x <= -f_lim, hi = x >= f_lim;
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
      for (int r = 0; r < 4; ++r) {
        const T x = cdt[CD_JAR + r];
        const bool on = x < 0;   // pyramid edge active
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
    bool done = scale * gnorm < tol || small_step;
    if (!done && iter > 0) {
      const T improvement = scale * (oldcost - cost);
      done = improvement < tol && (sizeof(T) == 8 || gnorm > T(0.5) * gnorm_prev);
    }
    gnorm_prev = gnorm;
    if (!done && iter >= maxiter) { status |= DEX_STATUS_SOLVER_MAXITER; done = true; }

    // ---- one factor / solve site for both the Newton direction and, once the iteration is over, the Euler step
    bool refactor;
    T diag, rhs;
    if (done) {
      // outputs of the forward pass, then (M + h B) qacc' = qfrc_smooth + qfrc_constraint (mj_Euler, implicit joint damping)
      qfrc_constraint = isdof ? (Ma - grad - fs) : T(0);  // = J^T f at the final iterate
      if (dbg.qacc && isdof) dbg.qacc[(size_t)env * nv + lane] = qacc;
      if (dbg.qfrc_smooth && isdof) dbg.qfrc_smooth[(size_t)env * nv + lane] = fs;
      if (dbg.qfrc_constraint && isdof) dbg.qfrc_constraint[(size_t)env * nv + lane] = qfrc_constraint;
      const bool badq = isdof && !(dx_abs(qacc) < T(1e10));  // mj_checkAcc (also catches NaN)
      if (dex_ballot(badq)) status |= DEX_STATUS_BADQACC;
      if (lane == 0) {
        st.niter[env] = iter - (zero_step ? 1 : 0);
        if (status) st.status[env] |= status;
      }
      if (!integrate) return;
      euler = true;
      if (structure != 0) { structure = 0; sp_init(L, m, 0); }
      refactor = true; diag = h * damping; rhs = isdof ? (fs + qfrc_constraint) : T(0);
    } else {
      const unsigned sig_changed = dex_ballot(asig != asig_prev || hdiag != hdiag_prev);
      asig_prev = asig; hdiag_prev = hdiag;
      refactor = iter == 0 || sig_changed;
      diag = hdiag; rhs = isdof ? -grad : T(0);
    }
    const DexSpConst& C = m.sp[structure];
    if (refactor) {
      sp_load(L, mrows, diag);
      if (!euler) {
        for (int c = 0; c < ncon; ++c) {
          const T* cdt = cdat + c * DEX_CDAT_STRIDE;
          const T w00 = cdt[CD_W], w01 = cdt[CD_W + 1], w02 = cdt[CD_W + 2], w11 = cdt[CD_W + 3], w22 = cdt[CD_W + 4];
          if (w00 == 0) continue;  // no active row (warp-uniform)
          const int dmax = *(const int*)(cdat + c * DEX_CDAT_STRIDE + CD_DMAX);
          const T a0 = Jc[(3 * c) * DEX_JC_ROW + lane], a1 = Jc[(3 * c + 1) * DEX_JC_ROW + lane], a2 = Jc[(3 * c + 2) * DEX_JC_ROW + lane];
          const T t0 = a0 * w00 + a1 * w01 + a2 * w02, t1 = a0 * w01 + a1 * w11, t2 = a0 * w02 + a2 * w22;
          sp_add_contact(L, a0, a1, a2, t0, t1, t2, dmax);
        }
      }
      int cfail = 0;
      sp_factor(L, C, rows, &cfail);
      if (cfail && !euler) status |= DEX_STATUS_CHOL_FAIL;
    }
    const T search = sp_solve(L, C, rows, rhs);
    if (euler) { qacc_e = search; break; }

    // ---- exact line search along `search` (PrimalSearch); see dex_dynamics_env
    vecB[lane] = isdof ? search : T(0);
    dex_syncwarp();
    const T Mv = dex_matvec(M, vecB, nv);
    dex_contact_dots(Jc, vecB, 3 * ncon, cdot);
    if (lane < ncon) {
      T* cdt = cdat + lane * DEX_CDAT_STRIDE;
      const T mu = cdt[CD_MU], sn = cdot[3 * lane], s1 = cdot[3 * lane + 1], s2 = cdot[3 * lane + 2];
      cdt[CD_JV + 0] = sn + mu * s1; cdt[CD_JV + 1] = sn - mu * s1;
      cdt[CD_JV + 2] = sn + mu * s2; cdt[CD_JV + 3] = sn - mu * s2;
    }
    dex_syncwarp();
    const T q1 = isdof ? search * (Ma - fs) : T(0), q2 = isdof ? search * Mv : T(0);
    const T g0 = dex_warp_sum(isdof ? search * grad : T(0));
    T alpha = g0 < 0 ? T(1) : T(0), lo = 0, hi = -1;   // not a descent direction: alpha stays 0
    for (int ls = 1; ls < 24 && g0 < 0; ++ls) {
      T g = q1 + alpha * q2, hh = q2;
      {
        const T x = qacc + alpha * search - f_aref;
        const bool on = f_loss > 0, lo = x <= -f_lim, hi = x >= f_lim;
        const T dg = lo ? -(f_loss * search) : (hi ? f_loss * search : f_D * x * search);
        g += on ? dg : T(0); hh += (on && !lo && !hi) ? f_D * search * search : T(0);
        const T jv = l_sign * search, xl = l_sign * qacc - l_aref + alpha * jv;
        const bool lon = (l_sign != 0) && (xl < 0);
        g += lon ? l_D * xl * jv : T(0); hh += lon ? l_D * jv * jv : T(0);
      }
      if (lane < ncon) {
        const T* cdt = cdat + lane * DEX_CDAT_STRIDE;
        const T D = cdt[CD_D];
        for (int r = 0; r < 4; ++r) {
          const T jv = cdt[CD_JV + r], x = cdt[CD_JAR + r] + alpha * jv;
          const bool on = x < 0;
          g += on ? D * x * jv : T(0); hh += on ? D * jv * jv : T(0);
        }
      }
      g = dex_warp_sum(g);
      hh = dex_warp_sum(hh);
      if (dx_abs(g) <= (sizeof(T) == 4 ? T(1e-3) : T(1e-12)) * dx_abs(g0)) break;
      if (g < 0) lo = alpha; else hi = alpha;
      T next = alpha - g / hh;
      if (hi >= 0 && (next <= lo || next >= hi)) next = T(0.5) * (lo + hi);
      const bool conv = dx_abs(next - alpha) <= (sizeof(T) == 4 ? T(1e-4) : T(1e-12)) * dx_abs(next);
      alpha = next;
      if (conv) break;
    }
    // a zero step ends the iteration like a converged one (the next pass finds done = true through small_step)
    zero_step = alpha == 0;
    small_step = zero_step || scale * (T(-0.5) * alpha * g0) < m.tolerance;
    // ---- move
    if (isdof) { qacc += alpha * search; Ma += alpha * Mv; }
    if (lane < ncon) {
      T* cdt = cdat + lane * DEX_CDAT_STRIDE;
      for (int r = 0; r < 4; ++r) cdt[CD_JAR + r] += alpha * cdt[CD_JV + r];
    }
    dex_syncwarp();
  }

  // ------------------------------------------------------------------ mj_Euler: integrate
  if (lane >= m.nv_euler) qacc_e = qacc;   // trees without joint damping: the solver's qacc is already the answer
  const T qv_new = qv + h * qacc_e;
  dex_syncwarp();
  qvel[lane] = qv_new;
  dex_syncwarp();
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

#endif  // DEX_SPARSE_CUH_
