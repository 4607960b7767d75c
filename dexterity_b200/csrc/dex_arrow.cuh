// dex_arrow.cuh — K2A: the constraint solver and integrator of dex_dynamics.cuh (SURVEY.md §8a rows 15, 19-20:
// mj_makeConstraint, mj_solNewton with exact line search, mj_Euler with implicit joint damping) for envs whose Newton
// Hessian is BLOCK-ARROWHEAD, with EIGHT LANES per env (four envs per warp).
//
// Structure.  The dofs of these scenes split into up to five BRANCHES (the fingers: subtrees the mass matrix does not
// couple to each other) and a BORDER of up to eight dofs (the wrist above the branching point, and the free object);
// DevModel::arrow_* holds the split (dex_host_model.h).  Friction-loss and joint-limit rows are diagonal.  A contact
// between a branch body and a border body (finger - cube, finger - palm), or between two border bodies (cube - palm),
// adds J^T W J inside one branch block, the border block and the coupling between the two.  So
//
//        | H_00            B_0^T |      H_aa (5 x 5) = M_aa + diag + sum_c Ja^T W Ja
//    H = |       ...        ...  |      B_a  (8 x 5) = M_oa        + sum_c Jo^T W Ja
//        | B_0   ...  B_4   H_oo |      H_oo (8 x 8) = M_oo + diag + sum_c Jo^T W Jo
//
// plus a low-rank term for every CROSS contact, one that joins two different branches (finger - finger): with U the contact's
// three frame rows and E its active pyramid edges, H = H0 + sum D (U^T e)(U^T e)^T, applied through the Woodbury identity in
// edge space: Z = H0^-1 U^T (three arrowhead solves per cross contact), G = U Z, S = diag(1 / D) + E^T G E (SPD, at most
// 2 x 4 edges, Cholesky with lane = edge), x = x0 - Z E S^-1 E^T U x0.  K1c (dex_compact_env) sends envs with more than
// DEX_AR_MAXCROSS cross contacts or more than DEX_AR_MAXCON contacts to the dense warp-per-env solver (dex_dynamics.cuh).
// The arrowhead factorisation is L_a = chol(H_aa) and P_a = B_a L_a^-T per branch (all branches at once, one lane each, in
// registers), the Schur complement S = H_oo - sum_a P_a P_a^T and its Cholesky factor (lane = border row, shuffles).
// Against the dense 30 x 30 factorisation with lane = row this cuts the sequential pivot chain from 30 to 5 + 8, and, because
// an env only needs eight lanes, every warp instruction serves four envs: 3-6 k warp instructions per env and mj_step
// instead of the dense kernel's 15 k.
//
// Two forms (template parameter CON of dex_arrow_warp): the full one, and the form for envs WITHOUT contacts, where all
// contact code drops out, the free object's acceleration is M^-1 qfrc_smooth and the Schur complement is 2 x 2.
//
// What it buys (B200, 16 384 envs, profiles/r02*): the contact-free form runs its half of the envs in 0.066 ms per launch.
// The full form needs 0.3 ms for the 44 % of envs it can take, no faster than the dense kernel on the same envs although
// it issues a third of the instructions: its unrolled Newton body (3 k SASS instructions, 48 KB) overflows the 32 KB
// instruction cache (26 % of the stall cycles wait for instructions), it needs ~240 registers (8 warps per SM; capping the
// registers spills and is slower), and the stage as a whole is bounded by its slowest env (19 contacts, 14 Newton
// iterations: 0.22 ms alone), which only the dense kernel takes.  The default therefore routes only the contact-free envs
// here (dex_set_solver mode 1); mode 2 adds the full form.
//
// Lane roles inside a group of eight (g = lane & 7), all views live on the same lanes at different times:
//   branch view  : lane g < 5 owns branch g (dof slots 0..4 of its per-dof register arrays)
//   border view  : lane g owns border row g (dof slot 5)
//   contact view : lane g owns contact g (per-contact scalars: impedance, the four pyramid rows)
// Cross-view data goes through the env's shared-memory slice; every array there is laid out [index][8] with the lane in
// the group fastest and the env slices DEX_AR_ENV_T = 8 (mod 32) words apart, so that the 32 lanes of a warp hit 32
// different banks whenever they use the same index.
//
// The arithmetic (rows, impedances, stopping rules, line search) is the dense kernel's, statement by statement; only the
// order of the floating-point sums differs.
#ifndef DEX_ARROW_CUH_
#define DEX_ARROW_CUH_

#include "dex_dynamics.cuh"

#define DEX_AR_G 8
#define DEX_AR_ENVS 4
#define DEX_AR_MAXCON 8      // contacts per env the structured solver takes (one lane per contact)
#define DEX_AR_MAXCROSS 2    // of which contacts that join two branches (handled as a low-rank correction)
#define DEX_AR_JC_CON 48     // contact-frame Jacobian of one contact in shared memory: [3 rows][16]: branch slots 0-4, border rows at 8-15

// shared memory per env (units of T)
enum {
  AR_PT = 0,                    // [25][8]  B^T -> P^T: packed branch column p = 5 a + k, border row fastest (aliased by the contact set-up records)
  AR_MAA = AR_PT + 25 * 8,      // [15][8]  lower triangle of M_aa of branch g
  AR_MAO = AR_MAA + 15 * 8,     // [10][8]  M[branch slot k][border row r < 2] at 2 k + r
  AR_MOO = AR_MAO + 10 * 8,     // [8][8]   M_oo[r'][r = g] (symmetric, full)
  AR_CF = AR_MOO + 8 * 8,       // [8][8]   per contact: frame-space force F(3), weights W00 W01 W02 W11 W22
  AR_LS = AR_CF,                //          aliased: factor of the Schur complement on its way from rows to columns, L[r = g][j] at j * 8 + g
  AR_XS = AR_CF + 64,           // [40]     packed vector exchange: branch entries at p = 5 a + k, border rows at 32 + r
  AR_QS = AR_XS + 40,           // [40]     qacc, packed (the row view reads it)
  AR_FR = AR_QS + 40,           // [40]     force of the dof's own rows (friction loss + limit), packed
  AR_HR = AR_FR + 40,           // [40]     their contribution to the diagonal of H, packed
  AR_CI = AR_HR + 40,           // [8]      per contact (int): its branch, -1 = border only
  AR_ENV_LITE_T = AR_CI + 8 + 16,   // = AR_JC: the contact-free form of the kernel (CON = false) stops here; 712 = 8 (mod 32)
  AR_JC = AR_CI + 8 + 16,       // [8 contacts][3 rows][16]  contact-frame Jacobians: first branch at 0-4; border rows, or the second branch of a
                                //          cross contact, at 8-15 (16-byte aligned)
  // low-rank part of the Hessian: contacts that join two branches (DEX_AR_MAXCROSS per env), see "cross contacts" below
  AR_Z = AR_JC + DEX_AR_MAXCON * DEX_AR_JC_CON,   // [6][40]  Z = H0^-1 U^T, one packed vector per frame row of a cross contact
  AR_G = AR_Z + 6 * 40,         // [6][8]   G = U Z
  AR_SL = AR_G + 48,            // [8][8]   Cholesky factor of the edge-space system, L[g][j] at j * 8 + g
  AR_SI = AR_SL + 64,           // [8]      1 / L[g][g]
  AR_TV = AR_SI + 8,            // [8]      U x0 going in, E y coming out (6 used)
  AR_CX = AR_TV + 8,            // [8][4]   per contact: mu, D, active-edge mask (int), -
  AR_XC = AR_CX + 32,           // [8]      int: contact index of cross slot 0, 1 (-1 = none)
  AR_ENV_T = AR_XC + 8 + 8      // 1512 = 8 (mod 32)
};
static_assert(AR_ENV_LITE_T % 32 == 8, "env slices must be 8 (mod 32) words apart");
static_assert(AR_ENV_T % 32 == 8, "env slices must be 8 (mod 32) words apart");
#define DEX_K2A_SMEM_T (DEX_AR_ENVS * AR_ENV_T)
#define DEX_K2L_SMEM_T (DEX_AR_ENVS * AR_ENV_LITE_T)

template <typename T> DEX_DEV T ar_gsum(T v) {   // sum over the group of eight
  v += dex_shfl_xor(v, 1); v += dex_shfl_xor(v, 2); v += dex_shfl_xor(v, 4);
  return v;
}
DEX_DEV bool ar_gany(bool p, int base) { return ((dex_ballot(p) >> base) & 0xffu) != 0; }

// index of (i, j), i >= j, in a packed lower triangle
#define AR_TRI(i, j) ((i) * ((i) + 1) / 2 + (j))

// ------------------------------------------------------------------------------------------------------------------
// Factorisation.  In: Ha = H_aa (packed lower, branch lanes), P^T area = B^T, Ho[j] = H_oo[g][j] (j <= g used).
// Out: Ha = L_a with Li = 1 / diag(L_a); P^T area = P^T; Ho = row g of L_S, Lc[j] = L_S[j][g] (column g), Si = 1 / L_S[g][g].
// NR = border rows that take part (8, or DevModel::arrow_nmo = 2 for the implicit-damping solve, where the rest is decoupled).
// ------------------------------------------------------------------------------------------------------------------
// Cholesky of an SPD matrix of order NR <= 8 spread over the group, lane g = row g (R[j], j <= g, are used): column sweep with
// shuffles.  Leaves row g of L in R and 1 / L[g][g] in inv.
template <typename T, int NR>
DEX_DEV void ar_chol_rows(int g, int base, T (&R)[8], T& inv, bool& bad) {
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    T d = dex_shfl(R[j], base | j);
    bad |= !(d > Real<T>::minval());
    d = dx_max(d, Real<T>::minval());
    const T iv = dx_rsqrt(d);
    const T l = R[j] * iv;   // L[g][j] for g >= j
    R[j] = l;
    inv = g == j ? iv : inv;
#pragma unroll
    for (int i = j + 1; i < NR; ++i) R[i] -= l * dex_shfl(l, base | i);
  }
}
// L L^T x = z with lane g holding row g (R) and column g (C) of L and inv = 1 / L[g][g]; z / x are lane g's entries.
template <typename T, int NR>
DEX_DEV T ar_tri_solve(int g, int base, const T (&R)[8], const T (&C)[8], T inv, T z) {
#pragma unroll
  for (int j = 0; j < NR; ++j) {
    const T yj = dex_shfl(z * inv, base | j);
    z -= (g > j ? R[j] : T(0)) * yj;
  }
  z *= inv;
#pragma unroll
  for (int j = NR - 1; j >= 0; --j) {
    const T xj = dex_shfl(z * inv, base | j);
    z -= (g < j ? C[j] : T(0)) * xj;
  }
  return z * inv;
}

template <typename T, int NR>
DEX_DEV void ar_factor(T* sm, int g, int base, T (&Ha)[15], T (&Li)[5], T (&Ho)[8], T (&Lc)[8], T& Si, bool& bad) {
  const bool isb = g < 5;
  // ---- branch blocks: Cholesky of the 5 x 5 in registers
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    T d = Ha[AR_TRI(j, j)];
#pragma unroll
    for (int t = 0; t < j; ++t) d -= Ha[AR_TRI(j, t)] * Ha[AR_TRI(j, t)];
    bad |= !(d > Real<T>::minval());
    d = dx_max(d, Real<T>::minval());
    const T inv = dx_rsqrt(d);
    Li[j] = inv;
    Ha[AR_TRI(j, j)] = d * inv;
#pragma unroll
    for (int i = j + 1; i < 5; ++i) {
      T s = Ha[AR_TRI(i, j)];
#pragma unroll
      for (int t = 0; t < j; ++t) s -= Ha[AR_TRI(i, t)] * Ha[AR_TRI(j, t)];
      Ha[AR_TRI(i, j)] = s * inv;
    }
  }
  // ---- P_a = B_a L_a^-T: one forward substitution per border row, rows in groups of four (16-byte accesses)
  if (isb) {
    T* pt = sm + AR_PT + (5 * g) * 8;
#pragma unroll
    for (int r0 = 0; r0 < NR; r0 += 4) {
      T v[5][4];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        if (NR - r0 >= 4) dex_ld4(pt + k * 8 + r0, v[k]);
        else { for (int q = 0; q < 4; ++q) v[k][q] = q < NR - r0 ? pt[k * 8 + r0 + q] : T(0); }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (r0 + q >= NR) continue;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          T s = v[k][q];
#pragma unroll
          for (int t = 0; t < k; ++t) s -= Ha[AR_TRI(k, t)] * v[t][q];
          v[k][q] = s * Li[k];
        }
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int q = 0; q < 4; ++q) if (r0 + q < NR) pt[k * 8 + r0 + q] = v[k][q];
      }
    }
  }
  dex_syncwarp();
  // ---- Schur complement, lane = border row: S[g][j] = H_oo[g][j] - sum_p P[g][p] P[j][p]
  {
    const T* pt = sm + AR_PT;
#pragma unroll 5
    for (int p = 0; p < 25; ++p) {
      const T mine = pt[p * 8 + (g < NR ? g : 0)];
      T row[8];
      if (NR == 8) { dex_ld4(pt + p * 8, row); dex_ld4(pt + p * 8 + 4, row + 4); }
      else { for (int j = 0; j < NR; ++j) row[j] = pt[p * 8 + j]; }
#pragma unroll
      for (int j = 0; j < NR; ++j) Ho[j] -= mine * row[j];
    }
  }
  ar_chol_rows<T, NR>(g, base, Ho, Si, bad);
  // column g of L_S through shared memory (the back substitution walks columns)
  {
    T* ls = sm + AR_LS;
#pragma unroll
    for (int j = 0; j < NR; ++j) ls[j * 8 + g] = Ho[j];
    dex_syncwarp();
#pragma unroll
    for (int j = 0; j < NR; ++j) Lc[j] = ls[(g < NR ? g : 0) * 8 + j];   // L[j][g], meaningful for j >= g
  }
}

// Solves H x = b with the factor above.  b / x: slots 0-4 = branch entries (branch lanes), slot 5 = border row g.
template <typename T, int NR>
DEX_DEV void ar_solve(T* sm, int g, int base, const T (&Ha)[15], const T (&Li)[5], const T (&Ho)[8], const T (&Lc)[8], T Si,
                      T (&x)[6]) {
  const bool isb = g < 5;
  T* xs = sm + AR_XS;
  // L_a y = b
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    T s = x[k];
#pragma unroll
    for (int t = 0; t < k; ++t) s -= Ha[AR_TRI(k, t)] * x[t];
    x[k] = s * Li[k];
  }
  if (isb) { for (int k = 0; k < 5; ++k) xs[5 * g + k] = x[k]; }
  dex_syncwarp();
  // z = b_o - sum_a P_a y_a (lane = border row)
  T z = g < NR ? x[5] : T(0);
  {
    const T* pt = sm + AR_PT;
#pragma unroll 5
    for (int p = 0; p < 25; ++p) z -= pt[p * 8 + (g < NR ? g : 0)] * xs[p];
    if (g >= NR) z = 0;
  }
  z = ar_tri_solve<T, NR>(g, base, Ho, Lc, Si, z);   // L_S w = z, L_S^T x_o = w
  if (g < NR) { x[5] = z; xs[32 + g] = z; }
  dex_syncwarp();
  // L_a^T x_a = y_a - P_a^T x_o
  if (isb) {
    const T* pt = sm + AR_PT + (5 * g) * 8;
    T xo[8];
#pragma unroll
    for (int r = 0; r < NR; ++r) xo[r] = xs[32 + r];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      T row[8];
      if (NR == 8) { dex_ld4(pt + k * 8, row); dex_ld4(pt + k * 8 + 4, row + 4); }
      else { for (int r = 0; r < NR; ++r) row[r] = pt[k * 8 + r]; }
      T s = x[k];
#pragma unroll
      for (int r = 0; r < NR; ++r) s -= row[r] * xo[r];
      x[k] = s;
    }
  }
#pragma unroll
  for (int k = 4; k >= 0; --k) {
    T s = x[k];
#pragma unroll
    for (int t = k + 1; t < 5; ++t) s -= Ha[AR_TRI(t, k)] * x[t];
    x[k] = s * Li[k];
  }
  dex_syncwarp();   // xs is reused by the caller
}

// y = M x for the arrowhead mass matrix (slots as in ar_solve).  Uses xs.
template <typename T>
DEX_DEV void ar_mulM(T* sm, int g, int nmo, const T (&x)[6], T (&y)[6]) {
  const bool isb = g < 5;
  T* xs = sm + AR_XS;
  xs[32 + g] = x[5];
  dex_syncwarp();
  const T* maa = sm + AR_MAA + g;
  const T* mao = sm + AR_MAO + g;
  const T xo0 = xs[32], xo1 = xs[33];
  T w0 = 0, w1 = 0;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    T s = 0;
#pragma unroll
    for (int t = 0; t < 5; ++t) s += maa[(t <= k ? AR_TRI(k, t) : AR_TRI(t, k)) * 8] * x[t];
    const T m0 = mao[(2 * k) * 8], m1 = mao[(2 * k + 1) * 8];
    y[k] = s + m0 * xo0 + m1 * xo1;
    w0 += m0 * x[k]; w1 += m1 * x[k];
  }
  if (!isb) { w0 = 0; w1 = 0; }
  w0 = ar_gsum(w0); w1 = ar_gsum(w1);
  T s = 0;
  const T* moo = sm + AR_MOO + g;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += moo[j * 8] * xs[32 + j];
  y[5] = s + (g == 0 ? w0 : (g == 1 ? w1 : T(0)));
  (void)nmo;
  dex_syncwarp();
}

// d[3] = Jc x for the lane's own contact (contact view); x is a packed vector in shared memory.  a = first branch (-1 = none),
// a2 = second branch of a cross contact (-1 = the columns 8-15 are the border rows).
template <typename T>
DEX_DEV void ar_contact_dot(const T* xs, const T* jc, int a, int a2, T* d) {
  T xa[5], xo[8];
#pragma unroll
  for (int k = 0; k < 5; ++k) xa[k] = a >= 0 ? xs[5 * a + k] : T(0);
#pragma unroll
  for (int r = 0; r < 8; ++r) xo[r] = a2 >= 0 ? (r < 5 ? xs[5 * a2 + r] : T(0)) : xs[32 + r];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    T ja[8], jo[8];
    dex_ld4(jc + q * 16, ja); ja[4] = jc[q * 16 + 4];
    dex_ld4(jc + q * 16 + 8, jo); dex_ld4(jc + q * 16 + 12, jo + 4);
    T s = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) s += ja[k] * xa[k];
#pragma unroll
    for (int r = 0; r < 8; ++r) s += jo[r] * xo[r];
    d[q] = s;
  }
}

// Joint-limit row of dof i (at most one side active): sign (+1 lower, -1 upper, 0 none), D = 1 / R and aref.
// The model tables come as a small struct of pointers: a reference to the DevModel kernel parameter would force a local copy.
template <typename T>
struct ArLimitTabs {
  const int *dof_jnt, *jnt_limited, *jnt_type, *jnt_qposadr;
  const T *jnt_margin, *jnt_range, *jnt_solimp, *jnt_solref, *dof_invweight0;
  T timestep;
};
template <typename T>
DEX_DEV_NOINLINE void ar_limit_row(const ArLimitTabs<T> m, int i, const T* qpos, T qv, T* lsign, T* lD, T* laref) {
  *lsign = 0; *lD = 0; *laref = 0;
  const int j = dex_ldg(m.dof_jnt + i);
  if (!dex_ldg(m.jnt_limited + j) || dex_ldg(m.jnt_type + j) < DEX_JNT_SLIDE) return;
  const T value = qpos[dex_ldg(m.jnt_qposadr + j)];
  const T margin = dex_ldg(m.jnt_margin + j);
  const T dlo = value - dex_ldg(m.jnt_range + 2 * j), dhi = dex_ldg(m.jnt_range + 2 * j + 1) - value;
  T pos = 0, sg = 0;
  if (dlo < margin) { sg = 1; pos = dlo; }
  else if (dhi < margin) { sg = -1; pos = dhi; }
  if (sg == 0) return;
  T si[5], sr[2], imp, K, B;
  for (int q = 0; q < 5; ++q) si[q] = dex_ldg(m.jnt_solimp + 5 * j + q);
  sr[0] = dex_ldg(m.jnt_solref + 2 * j); sr[1] = dex_ldg(m.jnt_solref + 2 * j + 1);
  dex_impedance(si, pos, margin, &imp);
  dex_kb(sr, si, m.timestep, &K, &B);
  const T R = dx_max(Real<T>::minval(), (1 - imp) / imp * dex_ldg(m.dof_invweight0 + i));
  *lsign = sg; *lD = 1 / R; *laref = -B * (sg * qv) - K * imp * (pos - margin);
}

// ------------------------------------------------------------------------------------------------------------------
// One warp = four envs (myenv of the lane's group, -1 = none).  smem: DEX_K2A_SMEM_T words.
// A fourth view, the ROW view: the dofs that own friction-loss / limit rows (DevModel::arrow_rowdof, at most 24) are dealt
// three to a lane, so that the row arithmetic of the force update and of every line-search evaluation runs on all eight
// lanes (in the dof view three lanes hold one dof and padding).  Rows read qacc / search from the packed arrays in shared
// memory and hand their force and Hessian diagonal back the same way.
// ------------------------------------------------------------------------------------------------------------------
// CON = false: the form for envs WITHOUT contacts (about half of them at any time).  All contact code drops out at compile
// time; the border rows beyond the leading arrow_nmo (the free object) carry no constraint at all, so their acceleration is
// M_oo^-1 qfrc_smooth, computed once, and the Newton iteration runs on the branches and the leading border rows only (a 2 x 2
// Schur complement).  The loop body then fits the 32 KB instruction cache, which the full form does not.
template <typename T, bool CON>
DEX_DEV void dex_arrow_warp(const DevModel<T>& m, const BatchState<T>& st, const DynDebug<T>& dbg, int myenv, T* smem, int integrate) {
  constexpr int NRN = CON ? 8 : 2;   // border rows of the Newton system
  const int lane = dex_lane(), g = lane & 7, base = lane & ~7;
  const bool live = myenv >= 0;
  const int env = live ? myenv : 0;   // dead groups compute on env 0 and write nothing
  T* sm = smem + (lane >> 3) * (CON ? AR_ENV_T : AR_ENV_LITE_T);
  const bool isb = g < 5;
  const T h = m.timestep;
  unsigned status = 0;

  // dof of every slot (-1 = identity padding)
  int dofs[6];
  {
    int t8[8];
    dex_ldg_int4(m.arrow_tab + 8 * g, t8); dex_ldg_int4(m.arrow_tab + 8 * g + 4, t8 + 4);
#pragma unroll
    for (int k = 0; k < 6; ++k) dofs[k] = t8[k];
  }

  const T* tree = st.tree + (size_t)env * DEX_TREE_ENV;
  const T* Mg = st.mscr + (size_t)env * DEX_MG_ENV;
  T* Jc = sm + AR_JC;
  const T* qposg = st.qpos + (size_t)env * st.qpos_stride;

  // ------------------------------------------------------------------ per-dof state and rows
  T qacc[6], fs[6], Ma[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int i = dofs[k];
    T qv = 0;
    qacc[k] = 0; fs[k] = 0;
    if (i >= 0) {
      qv = st.qvel[(size_t)env * st.qvel_stride + i];
      qacc[k] = st.warm[(size_t)env * st.warm_stride + i];
      fs[k] = tree[DEX_TREE_ROW_FS * 32 + i];
    }
    // packed qvel (for the contact-frame velocities below) and qacc; row outputs cleared
    const int p = k < 5 ? 5 * g + k : 32 + g;
    if (k == 5 || isb) { sm[AR_XS + p] = qv; sm[AR_QS + p] = qacc[k]; sm[AR_FR + p] = 0; sm[AR_HR + p] = 0; }
  }
  // rows of this lane (row view): friction loss {lim = R f, D = 1 / R, aref = -B qvel}, limit {sign, D, aref}
  const ArLimitTabs<T> ltab = {m.dof_jnt, m.jnt_limited, m.jnt_type, m.jnt_qposadr, m.jnt_margin, m.jnt_range, m.jnt_solimp, m.jnt_solref,
                               m.dof_invweight0, m.timestep};
  int rw_p[3];
  T rw_lim[3], rw_fD[3], rw_faref[3], rw_ls[3], rw_lD[3], rw_laref[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int i = dex_ldg(m.arrow_tab + DEX_AR_TAB_ROWS + g + 8 * q);
    rw_p[q] = -1; rw_lim[q] = Real<T>::big(); rw_fD[q] = 0; rw_faref[q] = 0; rw_ls[q] = 0; rw_lD[q] = 0; rw_laref[q] = 0;
    if (i >= 0) {
      rw_p[q] = dex_ldg(m.dof_pidx + i);
      const T qv = st.qvel[(size_t)env * st.qvel_stride + i];
      T fr[4];
      dex_ldg4(m.dof_flrow + 4 * i, fr);
      rw_lim[q] = fr[0]; rw_fD[q] = fr[1]; rw_faref[q] = -fr[2] * qv;
      ar_limit_row(ltab, i, qposg, qv, &rw_ls[q], &rw_lD[q], &rw_laref[q]);
    }
  }
  // ------------------------------------------------------------------ mass-matrix blocks
  {
    T* maa = sm + AR_MAA + g;
    T* mao = sm + AR_MAO + g;
    T* moo = sm + AR_MOO + g;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int i = dofs[k];
#pragma unroll
      for (int t = 0; t <= k; ++t) {
        const int j = dofs[t];
        maa[AR_TRI(k, t) * 8] = (i >= 0 && j >= 0) ? Mg[i * 32 + j] : (t == k ? T(1) : T(0));
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int j = r < m.arrow_nmo ? m.arrow_odof[r] : -1;
        mao[(2 * k + r) * 8] = (i >= 0 && j >= 0) ? Mg[i * 32 + j] : T(0);
      }
    }
    const int i = dofs[5];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int j = m.arrow_odof[r];
      moo[r * 8] = (i >= 0 && j >= 0) ? Mg[i * 32 + j] : (r == g ? T(1) : T(0));
    }
  }
  if (!CON) {
    // unconstrained border rows (beyond the leading arrow_nmo): qacc = M_oo^-1 qfrc_smooth, once (M_oo is block diagonal
    // between the two groups of rows, so the full 8 x 8 solve yields them; its leading rows are not used)
    T R8[8], C8[8], inv8 = 1;
    bool bad8 = false;
    const T* moo = sm + AR_MOO + g;
#pragma unroll
    for (int j = 0; j < 8; ++j) R8[j] = moo[j * 8];
    ar_chol_rows<T, 8>(g, base, R8, inv8, bad8);
    T* ls = sm + AR_LS;
#pragma unroll
    for (int j = 0; j < 8; ++j) ls[j * 8 + g] = R8[j];
    dex_syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) C8[j] = ls[g * 8 + j];
    const T a = ar_tri_solve<T, 8>(g, base, R8, C8, inv8, fs[5]);
    if (g >= m.arrow_nmo) { qacc[5] = a; sm[AR_QS + 32 + g] = a; }
    if (bad8) status |= DEX_STATUS_CHOL_FAIL;
    dex_syncwarp();
  }
  // ------------------------------------------------------------------ contacts (contact view: lane g = contact g)
  int ncon = (CON && live) ? st.ncon[env] : 0;
  if (ncon > DEX_AR_MAXCON) ncon = DEX_AR_MAXCON;   // K1c never sends such an env here
  const int ncmax = CON ? dex_warp_max(ncon) : 0;    // warp-uniform loop bound
  const bool hasc = CON && g < ncon;
  T c_mu = 0, c_D = 0, c_aref[4] = {0, 0, 0, 0}, c_kp = 0, c_B = 0;
  int c_a = -1, c_a2 = -1;   // branches of the contact: none / one (plus the border) / two (a cross contact)
  if (CON) {
    // set-up records in the (still unused) P^T area: [contact][16] = frame(9) pos(3) m1 m2 branch -
    T* rec = sm + AR_PT + g * 16;
    if (hasc) {
      const T* cc = st.contact + ((size_t)env * DEX_MAXCON + g) * DEX_CON_STRIDE;
      T c8[8];
      dex_ld4(cc, c8); dex_ld4(cc + 4, c8 + 4);
      T frame[9] = {c8[4], c8[5], c8[6], 0, 0, 0, 0, 0, 0};
      make_frame(frame);
      const int pr = (int)c8[7];
      const int b1 = dex_ldg(m.geom_body + dex_ldg(m.pair_geom1 + pr)), b2 = dex_ldg(m.geom_body + dex_ldg(m.pair_geom2 + pr));
      const unsigned m1 = (unsigned)dex_ldg(m.body_dofmask + b1), m2 = (unsigned)dex_ldg(m.body_dofmask + b2);
      const unsigned mx = m1 ^ m2;
#pragma unroll
      for (int a = 4; a >= 0; --a) if (mx & (unsigned)m.arrow_bmask[a]) { c_a2 = c_a; c_a = a; }   // c_a < c_a2
      for (int q = 0; q < 9; ++q) rec[q] = frame[q];
      rec[9] = c8[1]; rec[10] = c8[2]; rec[11] = c8[3];
      int* ri = reinterpret_cast<int*>(rec + 12);
      ri[0] = (int)m1; ri[1] = (int)m2; ri[2] = c_a; ri[3] = c_a2;
      // impedance, regulariser and the spring-damper of the four pyramid rows
      c_mu = dex_ldg(m.pair_friction + 5 * pr);
      const T cmargin = dex_ldg(m.pair_margin + pr) - dex_ldg(m.pair_gap + pr);
      T si[5], sr[2], imp, K;
      for (int q = 0; q < 5; ++q) si[q] = dex_ldg(m.pair_solimp + 5 * pr + q);
      sr[0] = dex_ldg(m.pair_solref + 2 * pr); sr[1] = dex_ldg(m.pair_solref + 2 * pr + 1);
      dex_impedance(si, c8[0], cmargin, &imp);
      dex_kb(sr, si, h, &K, &c_B);
      const T tran = dex_ldg(m.body_invweight0 + 2 * b1) + dex_ldg(m.body_invweight0 + 2 * b2);
      const T R0 = dx_max(Real<T>::minval(), (1 - imp) / imp * (tran + c_mu * c_mu * tran));
      const T Rpy = 2 * c_mu * c_mu * R0;
      c_D = 1 / Rpy;
      c_kp = K * imp * (c8[0] - cmargin);
      if (dbg.efc && live) {
        T* e = dbg.efc + ((size_t)env * DEX_MAXCON + g) * 8;
        e[0] = c_D; e[5] = c_mu; e[6] = imp; e[7] = Rpy;
      }
    }
    reinterpret_cast<int*>(sm + AR_CI)[g] = (c_a & 15) | ((c_a2 & 15) << 4);   // 15 = none
    sm[AR_CX + 4 * g] = c_mu; sm[AR_CX + 4 * g + 1] = c_D;
    reinterpret_cast<int*>(sm + AR_CX + 4 * g + 2)[0] = 0;
  }
  // cross contacts get the slots 0 .. nx-1 in contact order
  int nx = 0;
  if (CON) {
    const unsigned xm = (dex_ballot(hasc && c_a2 >= 0) >> base) & 0xffu;
    nx = dex_popc(xm);
    int* xc = reinterpret_cast<int*>(sm + AR_XC);
    if (g < DEX_AR_MAXCROSS) xc[g] = -1;
    dex_syncwarp();
    if (hasc && c_a2 >= 0) { const int rank = dex_popc(xm & ((1u << g) - 1u)); if (rank < DEX_AR_MAXCROSS) xc[rank] = g; }
    if (nx > DEX_AR_MAXCROSS) nx = DEX_AR_MAXCROSS;   // K1c never sends such an env here
  }
  const int nxmax = CON ? dex_warp_max(nx) : 0;   // warp-uniform: 0 = no env of this warp has a cross contact
  dex_syncwarp();
  // contact-frame Jacobians: branch lane = the 5 columns of its branch, every lane = border column g
  if (CON) {
    T cdo[6] = {0, 0, 0, 0, 0, 0}, refo[3] = {0, 0, 0};
    const int io = dofs[5];
    if (io >= 0) {
      for (int q = 0; q < 6; ++q) cdo[q] = tree[q * 32 + io];
      const int t = dex_ldg(m.dof_tree + io);
      for (int q = 0; q < 3; ++q) refo[q] = tree[DEX_TREE_ROW_REF * 32 + 3 * t + q];
    }
    for (int c = 0; c < ncmax; ++c) {
      if (c >= ncon) continue;
      const T* rec = sm + AR_PT + c * 16;
      T fr[12];
      dex_ld4(rec, fr); dex_ld4(rec + 4, fr + 4); dex_ld4(rec + 8, fr + 8);
      const int* ri = reinterpret_cast<const int*>(rec + 12);
      const unsigned m1 = (unsigned)ri[0], m2 = (unsigned)ri[1];
      const int ca = ri[2], ca2 = ri[3];
      T* jc = Jc + c * DEX_AR_JC_CON;
      {   // border column
        T j0 = 0, j1 = 0, j2 = 0;
        if (io >= 0) {
          const T sgn = T((m2 >> io) & 1u) - T((m1 >> io) & 1u);
          if (sgn != 0) {
            T off[3] = {fr[9] - refo[0], fr[10] - refo[1], fr[11] - refo[2]}, v[3];
            cross3(v, cdo, off);
            v[0] += cdo[3]; v[1] += cdo[4]; v[2] += cdo[5];
            j0 = sgn * dot3(fr, v); j1 = sgn * dot3(fr + 3, v); j2 = sgn * dot3(fr + 6, v);
          }
        }
        if (ca2 < 0) { jc[8 + g] = j0; jc[16 + 8 + g] = j1; jc[32 + 8 + g] = j2; }   // a cross contact has no border part: its columns 8-15 hold the second branch
      }
      if (isb) {   // branch columns (zero when the contact is not on this branch)
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int i = dofs[k];
          T j0 = 0, j1 = 0, j2 = 0;
          if ((ca == g || ca2 == g) && i >= 0) {
            const T sgn = T((m2 >> i) & 1u) - T((m1 >> i) & 1u);
            if (sgn != 0) {
              T cd[6];
              for (int q = 0; q < 6; ++q) cd[q] = tree[q * 32 + i];
              const int t = dex_ldg(m.dof_tree + i);
              T off[3] = {fr[9] - tree[DEX_TREE_ROW_REF * 32 + 3 * t], fr[10] - tree[DEX_TREE_ROW_REF * 32 + 3 * t + 1],
                          fr[11] - tree[DEX_TREE_ROW_REF * 32 + 3 * t + 2]}, v[3];
              cross3(v, cd, off);
              v[0] += cd[3]; v[1] += cd[4]; v[2] += cd[5];
              j0 = sgn * dot3(fr, v); j1 = sgn * dot3(fr + 3, v); j2 = sgn * dot3(fr + 6, v);
            }
          }
          if (ca == g || (ca < 0 && g == 0)) { jc[k] = j0; jc[16 + k] = j1; jc[32 + k] = j2; }   // border-only contacts: lane 0 writes the zeros
          if (ca2 == g) {
            jc[8 + k] = j0; jc[16 + 8 + k] = j1; jc[32 + 8 + k] = j2;
            if (k < 3) { jc[13 + k] = 0; jc[16 + 13 + k] = 0; jc[32 + 13 + k] = 0; }
          }
        }
      }
    }
  }
  dex_syncwarp();
  // reference accelerations of the pyramid rows from the contact-frame velocities, then jar = J qacc - aref
  T jar[4] = {0, 0, 0, 0};
  {
    T d[3] = {0, 0, 0};
    if (hasc) ar_contact_dot(sm + AR_XS, Jc + g * DEX_AR_JC_CON, c_a, c_a2, d);   // xs = qvel
    c_aref[0] = -c_B * (d[0] + c_mu * d[1]) - c_kp; c_aref[1] = -c_B * (d[0] - c_mu * d[1]) - c_kp;
    c_aref[2] = -c_B * (d[0] + c_mu * d[2]) - c_kp; c_aref[3] = -c_B * (d[0] - c_mu * d[2]) - c_kp;
    if (dbg.efc && live && hasc) {
      T* e = dbg.efc + ((size_t)env * DEX_MAXCON + g) * 8;
      e[1] = c_aref[0]; e[2] = c_aref[1]; e[3] = c_aref[2]; e[4] = c_aref[3];
    }
    dex_syncwarp();
#pragma unroll
    for (int k = 0; k < 5; ++k) if (isb) sm[AR_XS + 5 * g + k] = qacc[k];
    // ar_mulM stores the border entries of its argument in xs and leaves the branch entries alone
    ar_mulM(sm, g, m.arrow_nmo, qacc, Ma);
    if (hasc) {
      ar_contact_dot(sm + AR_XS, Jc + g * DEX_AR_JC_CON, c_a, c_a2, d);
      jar[0] = d[0] + c_mu * d[1] - c_aref[0]; jar[1] = d[0] - c_mu * d[1] - c_aref[1];
      jar[2] = d[0] + c_mu * d[2] - c_aref[2]; jar[3] = d[0] - c_mu * d[2] - c_aref[3];
    }
    dex_syncwarp();
  }

  // ------------------------------------------------------------------ Newton (mj_solNewton)
  const int nv = m.nv;
  const T scale = 1 / (m.meaninertia * (nv > 1 ? nv : 1));
  const T tol = dx_max(m.tolerance, sizeof(T) == 4 ? T(1e-6) : T(1e-12));
  T Ha[15], Li[5], Ho[8], Lc[8], Si = 1;
  T grad[6];
  T cost = 0, oldcost = 0, gnorm_prev = 0;
  int sig_prev = -1, iter = 0;
  bool small_step = false, done = !live, first = true;
  const int maxiter = m.iterations;
  for (;;) {
    // ---- forces, active sets, cost (PrimalUpdateConstraint)
    T mycost = 0;
    int sig = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) mycost += qacc[k] * (T(0.5) * Ma[k] - fs[k]);
#pragma unroll
    for (int q = 0; q < 3; ++q) {   // row view
      const int p = rw_p[q] >= 0 ? rw_p[q] : 0;
      const T a = sm[AR_QS + p];
      // friction loss: quadratic inside |x| < lim, linear outside, i.e. force -D clamp(x) and cost D (clamp(x) x - clamp(x)^2 / 2)
      const T x = a - rw_faref[q];
      const T xc = dx_min(dx_max(x, -rw_lim[q]), rw_lim[q]);
      const bool in = xc == x;
      mycost += rw_fD[q] * (xc * x - T(0.5) * xc * xc);
      T f = -(rw_fD[q] * xc), hd = in ? rw_fD[q] : T(0);
      sig |= in ? 1 << (2 * q) : 0;
      // limit: active while x < 0
      const T xl = rw_ls[q] * a - rw_laref[q];
      const bool on = xl < 0;
      const T xm = on ? xl : T(0);
      mycost += T(0.5) * rw_lD[q] * xm * xm;
      f -= rw_ls[q] * rw_lD[q] * xm;
      hd += on ? rw_lD[q] : T(0);
      sig |= (on && rw_lD[q] != 0) ? 2 << (2 * q) : 0;
      if (rw_p[q] >= 0) { sm[AR_FR + p] = f; sm[AR_HR + p] = hd; }
    }
    if (CON) {
      T f[4], w[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const T x = jar[r];
        const bool on = hasc && x < 0;   // pyramid edge active
        f[r] = on ? -c_D * x : T(0); w[r] = on ? c_D : T(0);
        mycost += on ? T(0.5) * c_D * x * x : T(0);
        sig |= on ? 1 << (8 + r) : 0;
      }
      reinterpret_cast<int*>(sm + AR_CX + 4 * g + 2)[0] = (sig >> 8) & 15;   // active pyramid edges (read by the cross-contact correction)
      T* cf = sm + AR_CF + g * 8;
      cf[0] = f[0] + f[1] + f[2] + f[3]; cf[1] = c_mu * (f[0] - f[1]); cf[2] = c_mu * (f[2] - f[3]);
      cf[3] = w[0] + w[1] + w[2] + w[3]; cf[4] = c_mu * (w[0] - w[1]); cf[5] = c_mu * (w[2] - w[3]);
      cf[6] = c_mu * c_mu * (w[0] + w[1]); cf[7] = c_mu * c_mu * (w[2] + w[3]);
    }
    dex_syncwarp();
    oldcost = cost;
    cost = ar_gsum(mycost);
    // ---- gradient = M a - qfrc_smooth - (row forces) - J^T f, contact by contact
#pragma unroll
    for (int k = 0; k < 6; ++k) grad[k] = Ma[k] - fs[k] - ((k == 5 || isb) ? sm[AR_FR + (k < 5 ? 5 * g + k : 32 + g)] : T(0));
    for (int c = 0; c < ncmax; ++c) {
      if (c >= ncon) continue;
      const T* cf = sm + AR_CF + c * 8;
      const T f0 = cf[0], f1 = cf[1], f2 = cf[2];
      const T* jc = Jc + c * DEX_AR_JC_CON;
      const int ci = reinterpret_cast<const int*>(sm + AR_CI)[c], ca = ci & 15, ca2 = ci >> 4;   // 15 = none
      if (ca2 == 15) grad[5] -= jc[8 + g] * f0 + jc[16 + 8 + g] * f1 + jc[32 + 8 + g] * f2;
      if (ca == g || ca2 == g) {
        const T* jb = jc + (ca == g ? 0 : 8);
        T a0[5], a1[5], a2[5];
        dex_ld4(jb, a0); a0[4] = jb[4]; dex_ld4(jb + 16, a1); a1[4] = jb[20]; dex_ld4(jb + 32, a2); a2[4] = jb[36];
#pragma unroll
        for (int k = 0; k < 5; ++k) grad[k] -= a0[k] * f0 + a1[k] * f1 + a2[k] * f2;
      }
    }
    if (!CON && g >= m.arrow_nmo) grad[5] = 0;   // unconstrained rows sit at their minimum already
    T g2 = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) g2 += grad[k] * grad[k];
    const T gnorm = dx_sqrt(ar_gsum(g2));
    if (!done) {
      if (scale * gnorm < tol || small_step) done = true;
      else if (iter > 0 && scale * (oldcost - cost) < tol && (sizeof(T) == 8 || gnorm > T(0.5) * gnorm_prev)) done = true;
      else if (iter >= maxiter) { status |= DEX_STATUS_SOLVER_MAXITER; done = true; }
      gnorm_prev = gnorm;
    }
    if (dex_ballot(!done) == 0) break;   // every env of the warp is finished

    // ---- Hessian and its factor; kept while the active set is unchanged.  When any env of the warp needs a new factor all
    // four rebuild theirs (same instructions either way; an unchanged active set reproduces the same factor bit for bit).
    const bool changed = ar_gany(sig != sig_prev, base) || first;
    sig_prev = sig;
    if (dex_ballot(changed && !done)) {
      first = false;
      {
        const T* maa = sm + AR_MAA + g;
#pragma unroll
        for (int q = 0; q < 15; ++q) Ha[q] = maa[q * 8];
#pragma unroll
        for (int k = 0; k < 5; ++k) Ha[AR_TRI(k, k)] += isb ? sm[AR_HR + 5 * g + k] : T(0);
        const T* moo = sm + AR_MOO + g;
        const T hdo = sm[AR_HR + 32 + g];
#pragma unroll
        for (int j = 0; j < 8; ++j) Ho[j] = moo[j * 8] + (j == g ? hdo : T(0));
        if (isb) {
          const T* mao = sm + AR_MAO + g;
          T* pt = sm + AR_PT + (5 * g) * 8;
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            pt[k * 8] = mao[(2 * k) * 8]; pt[k * 8 + 1] = mao[(2 * k + 1) * 8];
#pragma unroll
            for (int r = 2; r < 8; ++r) pt[k * 8 + r] = 0;
          }
        }
      }
      for (int c = 0; c < ncmax; ++c) {
        if (c >= ncon) continue;
        const T* cf = sm + AR_CF + c * 8;
        const T w00 = cf[3], w01 = cf[4], w02 = cf[5], w11 = cf[6], w22 = cf[7];
        if (w00 == 0) continue;   // no active edge (uniform within the group)
        const int ci = reinterpret_cast<const int*>(sm + AR_CI)[c];
        if ((ci >> 4) != 15) continue;   // a cross contact is not part of the arrowhead matrix: low-rank correction below
        const T* jc = Jc + c * DEX_AR_JC_CON;
        T o0[8], o1[8], o2[8];   // border columns of the three frame rows
        dex_ld4(jc + 8, o0); dex_ld4(jc + 12, o0 + 4); dex_ld4(jc + 24, o1); dex_ld4(jc + 28, o1 + 4);
        dex_ld4(jc + 40, o2); dex_ld4(jc + 44, o2 + 4);
        {   // border block, lane = row g
          const T b0 = jc[8 + g], b1 = jc[24 + g], b2 = jc[40 + g];
          const T t0 = b0 * w00 + b1 * w01 + b2 * w02, t1 = b0 * w01 + b1 * w11, t2 = b0 * w02 + b2 * w22;
#pragma unroll
          for (int j = 0; j < 8; ++j) Ho[j] += t0 * o0[j] + t1 * o1[j] + t2 * o2[j];
        }
        if ((ci & 15) == g) {   // branch block and coupling (only this lane touches these rows of P^T)
          T a0[5], a1[5], a2[5];
          dex_ld4(jc, a0); a0[4] = jc[4]; dex_ld4(jc + 16, a1); a1[4] = jc[20]; dex_ld4(jc + 32, a2); a2[4] = jc[36];
          T* pt = sm + AR_PT + (5 * g) * 8;
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const T t0 = a0[k] * w00 + a1[k] * w01 + a2[k] * w02, t1 = a0[k] * w01 + a1[k] * w11, t2 = a0[k] * w02 + a2[k] * w22;
#pragma unroll
            for (int t = 0; t <= k; ++t) Ha[AR_TRI(k, t)] += t0 * a0[t] + t1 * a1[t] + t2 * a2[t];
            T row[8];
            dex_ld4(pt + k * 8, row); dex_ld4(pt + k * 8 + 4, row + 4);
#pragma unroll
            for (int r = 0; r < 8; ++r) row[r] += t0 * o0[r] + t1 * o1[r] + t2 * o2[r];
#pragma unroll
            for (int r = 0; r < 8; ++r) pt[k * 8 + r] = row[r];
          }
        }
      }
      bool bad = false;
      ar_factor<T, NRN>(sm, g, base, Ha, Li, Ho, Lc, Si, bad);
      // ---- cross contacts: H = H0 + sum over their active pyramid edges D (U^T e)(U^T e)^T with U the contact's three frame
      // rows.  Woodbury in edge space: Z = H0^-1 U^T (three arrowhead solves per cross contact), G = U Z, and the SPD system
      // S = diag(1 / D) + E^T G E over the (at most 2 x 4) edges, factorised with lane = edge (lane 4 j + r: edge r of slot j).
      if (nxmax > 0) {
        const int* xc = reinterpret_cast<const int*>(sm + AR_XC);
        const int myslot = (hasc && c_a2 >= 0) ? (xc[0] == g ? 0 : 1) : -1;
#pragma unroll 1
        for (int i = 0; i < 3 * nxmax; ++i) {
          const int j = i / 3, q = i - 3 * j;
          const int c = xc[j];
          T rhs[6] = {0, 0, 0, 0, 0, 0};
          if (c >= 0) {
            const int ci = reinterpret_cast<const int*>(sm + AR_CI)[c];
            if ((ci & 15) == g || (ci >> 4) == g) {
              const T* jb = Jc + c * DEX_AR_JC_CON + q * 16 + ((ci & 15) == g ? 0 : 8);
#pragma unroll
              for (int k = 0; k < 5; ++k) rhs[k] = jb[k];
            }
          }
          ar_solve<T, 8>(sm, g, base, Ha, Li, Ho, Lc, Si, rhs);
          T* z = sm + AR_Z + i * 40;
#pragma unroll
          for (int k = 0; k < 6; ++k) if (k == 5 || isb) z[k < 5 ? 5 * g + k : 32 + g] = rhs[k];
          dex_syncwarp();
          if (myslot >= 0) {   // column i of G: the frame rows of this lane's cross contact against Z_i
            T d[3];
            ar_contact_dot(z, Jc + g * DEX_AR_JC_CON, c_a, c_a2, d);
#pragma unroll
            for (int qq = 0; qq < 3; ++qq) sm[AR_G + (3 * myslot + qq) * 8 + i] = d[qq];
          }
        }
        dex_syncwarp();
        // S row of this lane's edge
        const int ej = g >> 2, er = g & 3;
        const int ec = xc[ej];
        T mu[2];   // per slot: friction (the edge vectors (1, +-mu, 0), (1, 0, +-mu) are built from it)
        int act[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = xc[j];
          mu[j] = c >= 0 ? sm[AR_CX + 4 * c] : T(0);
          act[j] = c >= 0 ? reinterpret_cast<const int*>(sm + AR_CX + 4 * c + 2)[0] : 0;
        }
        const bool eon = ec >= 0 && (((ej == 0 ? act[0] : act[1]) >> er) & 1);
        const T em = ej == 0 ? mu[0] : mu[1];
        const T ev1 = er == 0 ? em : (er == 1 ? -em : T(0)), ev2 = er == 2 ? em : (er == 3 ? -em : T(0));
        T t6[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const T* gr = sm + AR_G + (3 * ej) * 8 + i;
          t6[i] = eon ? gr[0] + ev1 * gr[8] + ev2 * gr[16] : T(0);
        }
        T Se[8];
#pragma unroll
        for (int gg = 0; gg < 8; ++gg) {
          const int j2 = gg >> 2, r2 = gg & 3;
          const bool on2 = eon && ((act[j2] >> r2) & 1);
          const T m2 = mu[j2];
          const T v = t6[3 * j2] + (r2 == 0 ? m2 : (r2 == 1 ? -m2 : T(0))) * t6[3 * j2 + 1] + (r2 == 2 ? m2 : (r2 == 3 ? -m2 : T(0))) * t6[3 * j2 + 2];
          Se[gg] = on2 ? v : T(0);
        }
        {
          const T dinv = ec >= 0 ? T(1) / dx_max(sm[AR_CX + 4 * ec + 1], Real<T>::minval()) : T(1);
#pragma unroll
          for (int gg = 0; gg < 8; ++gg) if (gg == g) Se[gg] += eon ? dinv : T(1);   // an inactive edge keeps an identity row
        }
        T sinv = 1;
        ar_chol_rows<T, 8>(g, base, Se, sinv, bad);
#pragma unroll
        for (int j = 0; j < 8; ++j) sm[AR_SL + j * 8 + g] = Se[j];
        sm[AR_SI + g] = sinv;
        dex_syncwarp();
      }
      if (bad && !done) status |= DEX_STATUS_CHOL_FAIL;
    }
    // ---- search = -H^-1 grad
    T search[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) search[k] = -grad[k];
    ar_solve<T, NRN>(sm, g, base, Ha, Li, Ho, Lc, Si, search);
    if (!CON && g >= m.arrow_nmo) search[5] = 0;
    if (nxmax > 0) {   // x = x0 - Z E S^-1 E^T U x0
      const int* xc = reinterpret_cast<const int*>(sm + AR_XC);
#pragma unroll
      for (int k = 0; k < 6; ++k) if (k == 5 || isb) sm[AR_XS + (k < 5 ? 5 * g + k : 32 + g)] = search[k];
      dex_syncwarp();
      if (hasc && c_a2 >= 0) {
        T d[3];
        ar_contact_dot(sm + AR_XS, Jc + g * DEX_AR_JC_CON, c_a, c_a2, d);
        const int slot = xc[0] == g ? 0 : 1;
        sm[AR_TV + 3 * slot] = d[0]; sm[AR_TV + 3 * slot + 1] = d[1]; sm[AR_TV + 3 * slot + 2] = d[2];
      }
      dex_syncwarp();
      const int ej = g >> 2, er = g & 3;
      const int ec = xc[ej];
      const T em = ec >= 0 ? sm[AR_CX + 4 * ec] : T(0);
      const bool eon = ec >= 0 && ((reinterpret_cast<const int*>(sm + AR_CX + 4 * (ec >= 0 ? ec : 0) + 2)[0] >> er) & 1);
      const T ev1 = er == 0 ? em : (er == 1 ? -em : T(0)), ev2 = er == 2 ? em : (er == 3 ? -em : T(0));
      T r = eon ? sm[AR_TV + 3 * ej] + ev1 * sm[AR_TV + 3 * ej + 1] + ev2 * sm[AR_TV + 3 * ej + 2] : T(0);
      T SR[8], SC[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { SR[j] = sm[AR_SL + j * 8 + g]; SC[j] = sm[AR_SL + g * 8 + j]; }
      const T y = ar_tri_solve<T, 8>(g, base, SR, SC, sm[AR_SI + g], r);
      // E y: sum over the four edges of each slot
      T v0 = eon ? y : T(0), v1 = v0 * ev1, v2 = v0 * ev2;
      v0 += dex_shfl_xor(v0, 1); v1 += dex_shfl_xor(v1, 1); v2 += dex_shfl_xor(v2, 1);
      v0 += dex_shfl_xor(v0, 2); v1 += dex_shfl_xor(v1, 2); v2 += dex_shfl_xor(v2, 2);
      dex_syncwarp();
      if (er == 0) { sm[AR_TV + 3 * ej] = v0; sm[AR_TV + 3 * ej + 1] = v1; sm[AR_TV + 3 * ej + 2] = v2; }
      dex_syncwarp();
#pragma unroll 1
      for (int i = 0; i < 3 * nxmax; ++i) {
        const T vi = sm[AR_TV + i];
        const T* z = sm + AR_Z + i * 40;
#pragma unroll
        for (int k = 0; k < 6; ++k) if (k == 5 || isb) search[k] -= z[k < 5 ? 5 * g + k : 32 + g] * vi;
      }
      dex_syncwarp();
    }

    // ---- exact line search along `search` (PrimalSearch)
    T Mv[6];
#pragma unroll
    for (int k = 0; k < 5; ++k) if (isb) sm[AR_XS + 5 * g + k] = search[k];
    ar_mulM(sm, g, m.arrow_nmo, search, Mv);   // leaves xs = search (packed)
    T jv[4] = {0, 0, 0, 0};
    if (hasc) {
      T d[3];
      ar_contact_dot(sm + AR_XS, Jc + g * DEX_AR_JC_CON, c_a, c_a2, d);
      jv[0] = d[0] + c_mu * d[1]; jv[1] = d[0] - c_mu * d[1]; jv[2] = d[0] + c_mu * d[2]; jv[3] = d[0] - c_mu * d[2];
    }
    T q1 = 0, q2 = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) { q1 += search[k] * (Ma[k] - fs[k]); q2 += search[k] * Mv[k]; }
    // The first Newton step of the 1-D problem needs no row pass: at alpha = 0 the derivative is search . grad and the
    // curvature search . H search = -search . grad (H search = -grad), so the step lands on alpha = 1 exactly.
    T g0 = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) g0 += search[k] * grad[k];
    g0 = ar_gsum(g0);
    bool lsdone = done || !(g0 < 0);   // not a descent direction: alpha stays 0
    T alpha = lsdone ? T(0) : T(1), lo = 0, hi = -1;
    // rows of this lane along the search direction (xs = search, packed): x(alpha) = x0 + alpha s.  Friction loss contributes
    // A clamp(x) to the derivative and A s to the curvature inside the quadratic zone (A = D s); a limit B min(xl, 0) and B jv.
    T r_x0[3], r_s[3], r_A[3], r_l0[3], r_jv[3], r_B[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int p = rw_p[q] >= 0 ? rw_p[q] : 0;
      const T a = sm[AR_QS + p], sv = rw_p[q] >= 0 ? sm[AR_XS + p] : T(0);
      r_x0[q] = a - rw_faref[q]; r_s[q] = sv; r_A[q] = rw_fD[q] * sv;
      r_l0[q] = rw_ls[q] * a - rw_laref[q]; r_jv[q] = rw_ls[q] * sv; r_B[q] = rw_lD[q] * r_jv[q];
    }
    const T cDj[4] = {hasc ? c_D * jv[0] : T(0), hasc ? c_D * jv[1] : T(0), hasc ? c_D * jv[2] : T(0), hasc ? c_D * jv[3] : T(0)};
    q1 = ar_gsum(q1); q2 = ar_gsum(q2);
#pragma unroll 1
    for (int ls = 1; ls < 24; ++ls) {
      T gg = 0, hh = 0;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const T x = r_x0[q] + alpha * r_s[q];
        const T xc = dx_min(dx_max(x, -rw_lim[q]), rw_lim[q]);
        gg += r_A[q] * xc;
        hh += xc == x ? r_A[q] * r_s[q] : T(0);
        const T xl = r_l0[q] + alpha * r_jv[q];
        const bool lon = xl < 0;
        gg += lon ? r_B[q] * xl : T(0);
        hh += lon ? r_B[q] * r_jv[q] : T(0);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const T x = jar[r] + alpha * jv[r];
        const bool on = x < 0;
        gg += on ? cDj[r] * x : T(0); hh += on ? cDj[r] * jv[r] : T(0);
      }
      gg = ar_gsum(gg) + (q1 + alpha * q2);
      hh = ar_gsum(hh) + q2;
      if (!lsdone) {
        if (dx_abs(gg) <= (sizeof(T) == 4 ? T(1e-3) : T(1e-12)) * dx_abs(g0)) lsdone = true;
        if (!lsdone) {
          if (gg < 0) lo = alpha; else hi = alpha;
          T next = alpha - gg / hh;
          if (hi >= 0 && (next <= lo || next >= hi)) next = T(0.5) * (lo + hi);
          // converged in alpha: FP32 cannot push |g| below its rounding noise once g0 is small, and the outer Newton iteration
          // absorbs an error of this size in the step length
          if (dx_abs(next - alpha) <= (sizeof(T) == 4 ? T(1e-4) : T(1e-12)) * dx_abs(next)) lsdone = true;
          alpha = next;
        }
      }
      if (dex_ballot(!lsdone) == 0) break;
    }
    if (done) alpha = 0;
    if (!done) {
      if (alpha == 0) done = true;
      else {
        // MuJoCo's stopping rule "improvement * scale < tolerance", with the improvement taken from the line search
        // (-alpha g0 / 2 at an exact line minimum); FP32 cannot resolve it as a difference of costs
        small_step = scale * (T(-0.5) * alpha * g0) < m.tolerance;
        ++iter;
      }
    }
    // ---- move (alpha = 0 for finished envs)
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      qacc[k] += alpha * search[k]; Ma[k] += alpha * Mv[k];
      if (k == 5 || isb) sm[AR_QS + (k < 5 ? 5 * g + k : 32 + g)] = qacc[k];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) jar[r] += alpha * jv[r];
    dex_syncwarp();   // packed qacc is read by the row view of the next force update
  }

  // ------------------------------------------------------------------ outputs of the forward pass
  T fc[6];   // qfrc_constraint = J^T f at the final iterate
  bool badq = false;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    fc[k] = Ma[k] - grad[k] - fs[k];
    badq |= dofs[k] >= 0 && !(dx_abs(qacc[k]) < T(1e10));   // mj_checkAcc (also catches NaN)
    if (live && dofs[k] >= 0) {
      if (dbg.qacc) dbg.qacc[(size_t)env * nv + dofs[k]] = qacc[k];
      if (dbg.qfrc_smooth) dbg.qfrc_smooth[(size_t)env * nv + dofs[k]] = fs[k];
      if (dbg.qfrc_constraint) dbg.qfrc_constraint[(size_t)env * nv + dofs[k]] = fc[k];
    }
  }
  if (ar_gany(badq, base)) status |= DEX_STATUS_BADQACC;
  status = status | dex_shfl_xor(status, 1); status = status | dex_shfl_xor(status, 2); status = status | dex_shfl_xor(status, 4);
  if (live && g == 0) {
    st.niter[env] = iter;
    if (status) st.status[env] |= status;
  }
  if (!integrate) return;   // warp-uniform

  // ------------------------------------------------------------------ mj_Euler (implicit joint damping)
  // (M + h B) a = qfrc_smooth + qfrc_constraint over the branches and the leading border rows; the rest of the border
  // (the free object) is undamped and decoupled, the solver's qacc stands.
  T damp[6], qvn[6];
  {
    const T* maa = sm + AR_MAA + g;
#pragma unroll
    for (int q = 0; q < 15; ++q) Ha[q] = maa[q * 8];
#pragma unroll
    for (int k = 0; k < 6; ++k) damp[k] = dofs[k] >= 0 ? dex_ldg(m.dof_damping + dofs[k]) : T(0);
#pragma unroll
    for (int k = 0; k < 5; ++k) Ha[AR_TRI(k, k)] += h * damp[k];
    const T* moo = sm + AR_MOO + g;
#pragma unroll
    for (int j = 0; j < 8; ++j) Ho[j] = moo[j * 8] + (j == g ? h * damp[5] : T(0));
    if (isb) {
      const T* mao = sm + AR_MAO + g;
      T* pt = sm + AR_PT + (5 * g) * 8;
#pragma unroll
      for (int k = 0; k < 5; ++k) { pt[k * 8] = mao[(2 * k) * 8]; pt[k * 8 + 1] = mao[(2 * k + 1) * 8]; }
    }
    dex_syncwarp();
    bool bad = false;
    ar_factor<T, 2>(sm, g, base, Ha, Li, Ho, Lc, Si, bad);
    T rhs[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) rhs[k] = fs[k] + fc[k];
    ar_solve<T, 2>(sm, g, base, Ha, Li, Ho, Lc, Si, rhs);
    if (g >= m.arrow_nmo) rhs[5] = qacc[5];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int i = dofs[k];
      const T qv = i >= 0 ? st.qvel[(size_t)env * st.qvel_stride + i] : T(0);
      qvn[k] = qv + h * rhs[k];
      if (k < 5) { if (isb) sm[AR_XS + 5 * g + k] = qvn[k]; } else sm[AR_XS + 32 + g] = qvn[k];
    }
  }
  // new qvel (xs) and the solver's qacc (AR_QS) are packed in shared memory: the write-back is a rolled loop
  dex_syncwarp();   // AR_QS holds the solver's qacc since the last move
  if (live) {
    const T* xs = sm + AR_XS;
    const T* qs = sm + AR_QS;
    T* qpe = st.qpos + (size_t)env * st.qpos_stride;
#pragma unroll 1
    for (int k = 0; k < 6; ++k) {
      const int i = dex_ldg(m.arrow_tab + 8 * g + k);
      if (i < 0) continue;
      const int j = dex_ldg(m.dof_jnt + i), p = k < 5 ? 5 * g + k : 32 + g;
      if (dex_ldg(m.jnt_type + j) != DEX_JNT_FREE) { const int qa = dex_ldg(m.jnt_qposadr + j); qpe[qa] = qpe[qa] + h * xs[p]; }
      st.qvel[(size_t)env * st.qvel_stride + i] = xs[p];
      st.warm[(size_t)env * st.warm_stride + i] = qs[p];
    }
    const int io = dofs[5];
    if (io >= 0) {
      const int j = dex_ldg(m.dof_jnt + io), da = dex_ldg(m.jnt_dofadr + j);
      if (dex_ldg(m.jnt_type + j) == DEX_JNT_FREE && io == da) {   // the lane of a free joint's first dof integrates its pose
        T* qp = qpe + dex_ldg(m.jnt_qposadr + j);
        T v6[6];
        for (int q = 0; q < 6; ++q) v6[q] = xs[dex_ldg(m.dof_pidx + da + q)];
        for (int q = 0; q < 3; ++q) qp[q] = qp[q] + h * v6[q];
        T w[3] = {v6[3], v6[4], v6[5]};
        T q4[4] = {qp[3], qp[4], qp[5], qp[6]};
        const T n = dx_sqrt(dot3(w, w));
        if (n >= Real<T>::minval()) {  // mju_quatIntegrate
          T ax[3] = {w[0] / n, w[1] / n, w[2] / n}, dq[4], r[4];
          axisangle2quat(dq, ax, h * n);
          quat_mul(r, q4, dq);
          quat_normalize(r);
          q4[0] = r[0]; q4[1] = r[1]; q4[2] = r[2]; q4[3] = r[3];
        }
        for (int q = 0; q < 4; ++q) qp[3 + q] = q4[q];
      }
    }
    if (g == 0) {
      if (st.time64) { const double t = st.time64[2 * (size_t)env] + m.timestep_d; st.time64[2 * (size_t)env] = t; st.time[env] = (T)t; }
      else st.time[env] += h;
    }
  }
}

#endif  // DEX_ARROW_CUH_
