// dex_host_model.h — host-side preparation of the device model from the compiled blob
// (include/dex_model_blob.h).  Pure C++ (no CUDA): derives the tables the kernels want
// (tree depths, per-body dof masks, body-frame inertia matrices and hull vertices, body
// bounding spheres, candidate pairs grouped by body pair, per-tree reference bodies) and
// validates the model against the kernels' compile-time capacities.
#ifndef DEX_HOST_MODEL_H_
#define DEX_HOST_MODEL_H_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/dex_model_blob.h"
#include "dex_model.cuh"

struct HostModel {
  int nbody, njnt, nq, nv, nu, ntendon, ngeom, npair, nsite, nbodypair, ntree, maxdepth;
  int nconmax, iterations, mpr_iterations, nv_euler;
  int k0_rounds, chol_split;
  int arrow_ok, arrow_nmo, arrow_bdof[5][5], arrow_odof[8], arrow_bmask[5], arrow_rowdof[24], arrow_omask;
  int sp_ok;
  DexSpConst sp[2];
  double timestep, gravity[3], tolerance, mpr_tolerance, meaninertia, max_margin;
#define DEX_X_I(name) std::vector<int> name;
#define DEX_X_R(name) std::vector<double> name;
#define DEX_X(t, name) DEX_X_##t(name)
  DEX_DEVMODEL_ARRAYS(DEX_X)
#undef DEX_X
#undef DEX_X_I
#undef DEX_X_R
};

namespace dex_host {

inline void q2m(const double* q, double* m) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}

// returns "" on success, else an error message
inline std::string build(const void* blob, uint64_t nbytes, HostModel& M) {
  const uint8_t* base = (const uint8_t*)blob;
  const DexBlobHeader* hd = (const DexBlobHeader*)blob;
  if (nbytes < sizeof(DexBlobHeader) || hd->magic != DEX_BLOB_MAGIC) return "bad blob magic";
  if (hd->version != DEX_BLOB_VERSION) return "blob version mismatch";
  if (hd->nfields != DEXF_COUNT || hd->total_bytes != nbytes) return "blob size/field count mismatch";
  const DexBlobEntry* ent = (const DexBlobEntry*)(base + sizeof(DexBlobHeader));
  auto I = [&](int f) { return std::vector<int>((const int32_t*)(base + ent[f].offset), (const int32_t*)(base + ent[f].offset) + ent[f].count); };
  auto F = [&](int f) { return std::vector<double>((const double*)(base + ent[f].offset), (const double*)(base + ent[f].offset) + ent[f].count); };
  std::vector<int> D = I(DEXF_DIMS);
  std::vector<double> O = F(DEXF_OPT);
  M.nbody = D[DEXD_NBODY]; M.njnt = D[DEXD_NJNT]; M.nq = D[DEXD_NQ]; M.nv = D[DEXD_NV]; M.nu = D[DEXD_NU];
  M.ntendon = D[DEXD_NTENDON]; M.ngeom = D[DEXD_NGEOM]; M.npair = D[DEXD_NPAIR]; M.nsite = D[DEXD_NSITE];
  M.nconmax = D[DEXD_NCONMAX]; M.iterations = D[DEXD_SOLVER_ITERATIONS]; M.mpr_iterations = D[DEXD_MPR_ITERATIONS];
  M.timestep = O[DEXO_TIMESTEP]; M.gravity[0] = O[DEXO_GRAVITY_X]; M.gravity[1] = O[DEXO_GRAVITY_Y]; M.gravity[2] = O[DEXO_GRAVITY_Z];
  M.tolerance = O[DEXO_TOLERANCE]; M.mpr_tolerance = O[DEXO_MPR_TOLERANCE]; M.meaninertia = O[DEXO_MEANINERTIA];
  if (M.nbody > DEX_MAXBODY) return "nbody exceeds DEX_MAXBODY";
  if (M.nv > DEX_MAXDOF) return "nv exceeds DEX_MAXDOF (32)";
  if (M.njnt > DEX_MAXJNT) return "njnt exceeds DEX_MAXJNT";
  if (M.nu > DEX_MAXU) return "nu exceeds DEX_MAXU";
  if (M.nq > 48) return "nq exceeds 48";
  if (M.npair >= (1 << 24)) return "too many candidate pairs";

  M.body_parent = I(DEXF_BODY_PARENT); M.body_jntadr = I(DEXF_BODY_JNTADR); M.body_jntnum = I(DEXF_BODY_JNTNUM);
  M.body_dofadr = I(DEXF_BODY_DOFADR); M.body_dofnum = I(DEXF_BODY_DOFNUM); M.body_weld = I(DEXF_BODY_WELD);
  M.body_pos = F(DEXF_BODY_POS); M.body_quat = F(DEXF_BODY_QUAT); M.body_ipos = F(DEXF_BODY_IPOS);
  M.body_mass = F(DEXF_BODY_MASS); M.body_xfrc = F(DEXF_BODY_XFRC); M.body_invweight0 = F(DEXF_BODY_INVWEIGHT0);
  M.jnt_type = I(DEXF_JNT_TYPE); M.jnt_body = I(DEXF_JNT_BODY); M.jnt_qposadr = I(DEXF_JNT_QPOSADR);
  M.jnt_dofadr = I(DEXF_JNT_DOFADR); M.jnt_limited = I(DEXF_JNT_LIMITED); M.jnt_pos = F(DEXF_JNT_POS);
  M.jnt_axis = F(DEXF_JNT_AXIS); M.jnt_range = F(DEXF_JNT_RANGE); M.jnt_margin = F(DEXF_JNT_MARGIN);
  M.jnt_solref = F(DEXF_JNT_SOLREF); M.jnt_solimp = F(DEXF_JNT_SOLIMP); M.qpos0 = F(DEXF_QPOS0);
  M.dof_body = I(DEXF_DOF_BODY); M.dof_jnt = I(DEXF_DOF_JNT); M.dof_parent = I(DEXF_DOF_PARENT);
  M.dof_armature = F(DEXF_DOF_ARMATURE); M.dof_damping = F(DEXF_DOF_DAMPING); M.dof_frictionloss = F(DEXF_DOF_FRICTIONLOSS);
  M.dof_invweight0 = F(DEXF_DOF_INVWEIGHT0); M.dof_solref = F(DEXF_DOF_SOLREF); M.dof_solimp = F(DEXF_DOF_SOLIMP);
  M.ten_adr = I(DEXF_TEN_ADR); M.ten_num = I(DEXF_TEN_NUM); M.ten_limited = I(DEXF_TEN_LIMITED);
  M.ten_range = F(DEXF_TEN_RANGE); M.ten_margin = F(DEXF_TEN_MARGIN); M.ten_invweight0 = F(DEXF_TEN_INVWEIGHT0);
  M.ten_solref = F(DEXF_TEN_SOLREF); M.ten_solimp = F(DEXF_TEN_SOLIMP); M.wrap_jnt = I(DEXF_WRAP_JNT); M.wrap_coef = F(DEXF_WRAP_COEF);
  M.act_trntype = I(DEXF_ACT_TRNTYPE); M.act_trnid = I(DEXF_ACT_TRNID); M.act_ctrllimited = I(DEXF_ACT_CTRLLIMITED);
  M.act_forcelimited = I(DEXF_ACT_FORCELIMITED); M.act_gear = F(DEXF_ACT_GEAR); M.act_gainprm = F(DEXF_ACT_GAINPRM);
  M.act_biasprm = F(DEXF_ACT_BIASPRM); M.act_ctrlrange = F(DEXF_ACT_CTRLRANGE); M.act_forcerange = F(DEXF_ACT_FORCERANGE);
  M.geom_type = I(DEXF_GEOM_TYPE); M.geom_body = I(DEXF_GEOM_BODY); M.geom_pos = F(DEXF_GEOM_POS);
  M.geom_size = F(DEXF_GEOM_SIZE); M.geom_rbound = F(DEXF_GEOM_RBOUND);
  M.site_body = I(DEXF_SITE_BODY); M.site_pos = F(DEXF_SITE_POS);
  for (int t = 0; t < M.ntendon; ++t) if (M.ten_limited[t]) return "limited tendons are not supported yet";
  if (M.ntendon > DEX_MAXTEN) return "ntendon exceeds DEX_MAXTEN";

  // ---- tree structure
  const int nb = M.nbody, nv = M.nv;
  M.body_depth.assign(nb, 0);
  M.maxdepth = 0;
  for (int b = 1; b < nb; ++b) {
    if (M.body_parent[b] < 0 || M.body_parent[b] >= b) return "bodies must be ordered parent before child";
    M.body_depth[b] = M.body_depth[M.body_parent[b]] + 1; M.maxdepth = std::max(M.maxdepth, M.body_depth[b]);
  }
  if (M.maxdepth >= 12) return "kinematic tree deeper than DEX_MAXDEPTH (12)";
  {  // the tree kernel walks the bodies with a depth-indexed stack: requires DFS pre-order (MuJoCo's body order)
    std::vector<int> open(M.maxdepth + 1, 0);
    for (int b = 1; b < nb; ++b) {
      if (open[M.body_depth[b] - 1] != M.body_parent[b]) return "bodies are not in depth-first pre-order";
      open[M.body_depth[b]] = b;
    }
  }
  // ---- schedule of the 8-lanes-per-env tree kernel (K0): round = tree depth, a lane follows a chain of first
  //      children.  Per (round, lane) two 4-int records: {body, parent's lane, joints adr|num<<16, tree id} and
  //      {body, dofs adr|num<<16, children adr|num<<16, 0}; per dof its ancestor chain (self first, -1 terminated).
  //      (filled in below once the tree ids are known)
  std::vector<int> k0_body, k0_src;
  {
    const int G = 8;
    M.k0_rounds = M.maxdepth;
    k0_body.assign(std::max(1, M.maxdepth * G), -1);
    k0_src.assign(std::max(1, M.maxdepth * G), -1);
    std::vector<int> lane_of(nb, -1);
    for (int d = 1; d <= M.maxdepth && M.k0_rounds; ++d) {
      int* rb = &k0_body[(d - 1) * G];
      int* rs = &k0_src[(d - 1) * G];
      for (int pass = 0; pass < 2 && M.k0_rounds; ++pass)   // pass 0: first children keep the parent's lane
        for (int b = 1; b < nb; ++b) {
          if (M.body_depth[b] != d || lane_of[b] >= 0) continue;
          const int p = M.body_parent[b];
          int l = -1;
          if (pass == 0) { if (p > 0 && rb[lane_of[p]] < 0) l = lane_of[p]; else continue; }
          else for (int k = 0; k < G; ++k) if (rb[k] < 0) { l = k; break; }
          if (l < 0) { M.k0_rounds = 0; break; }   // level wider than 8 bodies
          rb[l] = b; rs[l] = p > 0 ? lane_of[p] : -1; lane_of[b] = l;
        }
    }
    M.dof_anc8.assign(std::max(1, 8 * nv), -1);
    for (int i = 0; i < nv; ++i) {
      int t = 0;
      for (int j = i; j >= 0; j = M.dof_parent[j]) { if (t == 8) { M.k0_rounds = 0; break; } M.dof_anc8[8 * i + t++] = j; }
    }
  }
  std::vector<int> broot = I(DEXF_BODY_ROOT);
  std::map<int, int> tree_of_root;
  M.body_tree.assign(nb, 0);
  tree_of_root[0] = 0;  // world
  for (int b = 1; b < nb; ++b) {
    if (!tree_of_root.count(broot[b])) { int t = (int)tree_of_root.size(); tree_of_root[broot[b]] = t; }
    M.body_tree[b] = tree_of_root[broot[b]];
  }
  M.ntree = (int)tree_of_root.size();
  {
    std::vector<int> cadr(nb, 0), cnum(nb, 0);
    M.body_child.clear();
    for (int b = 1; b < nb; ++b) {   // children lists (the world's are not needed: nothing accumulates into the world body)
      cadr[b] = (int)M.body_child.size();
      for (int c = b + 1; c < nb; ++c) if (M.body_parent[c] == b) { M.body_child.push_back(c); cnum[b]++; }
    }
    if (M.body_child.empty()) M.body_child.assign(1, 0);
    M.k0_rec.assign(k0_body.size() * 8, 0);
    for (size_t e = 0; e < k0_body.size(); ++e) {
      const int b = k0_body[e];
      int* r = &M.k0_rec[8 * e];
      r[0] = b; r[1] = k0_src[e]; r[4] = b;
      if (b >= 0) {
        r[2] = M.body_jntadr[b] | (M.body_jntnum[b] << 16); r[3] = M.body_tree[b];
        r[5] = M.body_dofadr[b] | (M.body_dofnum[b] << 16); r[6] = cadr[b] | (cnum[b] << 16);
        if (M.body_dofnum[b] == 0) r[5] = 0;
        if (M.body_jntnum[b] == 0) r[2] = 0;
      }
    }
  }
  if (M.ntree > 5) return "more than 5 kinematic trees";
  M.dof_tree.assign(nv, 0);
  for (int i = 0; i < nv; ++i) M.dof_tree[i] = M.body_tree[M.dof_body[i]];
  {  // dofs of trees that carry joint damping must be a prefix for the reduced implicit-Euler solve
    std::vector<int> tree_damped(M.ntree, 0);
    for (int i = 0; i < nv; ++i) if (M.dof_damping[i] > 0) tree_damped[M.dof_tree[i]] = 1;
    M.nv_euler = 0;
    for (int i = 0; i < nv; ++i) if (tree_damped[M.dof_tree[i]]) M.nv_euler = i + 1;
  }
  {  // largest multiple of 4 at which the dofs split into two groups of whole kinematic trees (M is block diagonal there)
    M.chol_split = nv;
    for (int s4 = 4; s4 < nv; s4 += 4) if (M.dof_tree[s4] != M.dof_tree[s4 - 1]) M.chol_split = s4;
  }
  // ---- arrowhead decomposition for the structured solver (dex_arrow.cuh).  Border = every dof with two or more child
  //      dofs plus its ancestors (the rows above the branching point), followed by the dofs of trees without branching
  //      (a free object); the connected pieces that remain are the branches.  M couples a branch only to itself and to
  //      its border ancestors, so H = M + diag + J^T W J is block-arrowhead as long as no contact joins two branches.
  {
    M.arrow_ok = 1; M.arrow_nmo = 0;
    for (int a = 0; a < 5; ++a) { M.arrow_bmask[a] = 0; for (int k = 0; k < 5; ++k) M.arrow_bdof[a][k] = -1; }
    for (int r = 0; r < 8; ++r) M.arrow_odof[r] = -1;
    std::vector<int> nchild(nv, 0), isborder(nv, 0), tree_branches(M.ntree, 0);
    for (int i = 0; i < nv; ++i) if (M.dof_parent[i] >= 0) nchild[M.dof_parent[i]]++;
    for (int i = 0; i < nv; ++i) if (nchild[i] >= 2) { tree_branches[M.dof_tree[i]] = 1; for (int j = i; j >= 0; j = M.dof_parent[j]) isborder[j] = 1; }
    int nbo = 0;
    for (int i = 0; i < nv; ++i) if (isborder[i]) { if (nbo < 8) M.arrow_odof[nbo] = i; nbo++; }
    M.arrow_nmo = nbo;
    for (int i = 0; i < nv; ++i) if (!tree_branches[M.dof_tree[i]]) { isborder[i] = 1; if (nbo < 8) M.arrow_odof[nbo] = i; nbo++; }
    if (nbo > 8 || M.arrow_nmo > 2 || M.arrow_nmo == 0) M.arrow_ok = 0;
    int nbr = 0;
    std::vector<int> branch_of(nv, -1);
    for (int i = 0; i < nv && M.arrow_ok; ++i) {
      if (isborder[i]) continue;
      const int p = M.dof_parent[i];
      int a = (p >= 0 && !isborder[p]) ? branch_of[p] : -1;
      if (a < 0) { if (nbr == 5) { M.arrow_ok = 0; break; } a = nbr++; }
      branch_of[i] = a;
      int k = 0;
      while (k < 5 && M.arrow_bdof[a][k] >= 0) ++k;
      if (k == 5) { M.arrow_ok = 0; break; }
      M.arrow_bdof[a][k] = i; M.arrow_bmask[a] |= 1 << i;
    }
    // the reduced implicit-Euler solve takes the damped trees = the branches and the leading border rows
    for (int i = 0; i < nv && M.arrow_ok; ++i) {
      bool lead = !isborder[i];
      for (int r = 0; r < M.arrow_nmo; ++r) if (M.arrow_odof[r] == i) lead = true;
      if (lead != (i < M.nv_euler)) M.arrow_ok = 0;
    }
    {
      int nrow = 0;
      for (int r = 0; r < 24; ++r) M.arrow_rowdof[r] = -1;
      for (int i = 0; i < nv; ++i) {
        const int j = M.dof_jnt[i];
        const bool limited = M.jnt_limited[j] && M.jnt_type[j] >= DEX_JNT_SLIDE;
        if (!(M.dof_frictionloss[i] > 0) && !limited) continue;
        if (nrow < 24) M.arrow_rowdof[nrow] = i;
        nrow++;
      }
      if (nrow > 24) M.arrow_ok = 0;
    }
    M.arrow_omask = 0;
    for (int r = 0; r < 8; ++r) if (M.arrow_odof[r] >= 0) M.arrow_omask |= 1 << M.arrow_odof[r];
    M.arrow_tab.assign(64 + 24, -1);
    for (int g = 0; g < 8; ++g) {
      for (int k = 0; k < 5; ++k) if (g < 5) M.arrow_tab[8 * g + k] = M.arrow_bdof[g][k];
      M.arrow_tab[8 * g + 5] = M.arrow_odof[g];
    }
    for (int r = 0; r < 24; ++r) M.arrow_tab[64 + r] = M.arrow_rowdof[r];
    M.dof_pidx.assign(std::max(1, nv), 0);   // packed position of a dof: 5 * branch + slot, or 32 + border row
    if (M.arrow_ok) {
      for (int a = 0; a < 5; ++a) for (int k = 0; k < 5; ++k) if (M.arrow_bdof[a][k] >= 0) M.dof_pidx[M.arrow_bdof[a][k]] = 5 * a + k;
      for (int r = 0; r < 8; ++r) if (M.arrow_odof[r] >= 0) M.dof_pidx[M.arrow_odof[r]] = 32 + r;
    }
    // dof friction-loss rows are state independent up to aref = -B qvel (mj_makeImpedance at pos = margin = 0):
    // per dof {R * frictionloss, D = 1 / R, B, 0}; {huge, 0, 0, 0} without friction loss
    M.dof_flrow.assign(std::max(1, 4 * nv), 0.0);
    for (int i = 0; i < nv; ++i) {
      double* o = &M.dof_flrow[4 * i];
      o[0] = 1e30; o[1] = 0; o[2] = 0;
      const double fl = M.dof_frictionloss[i];
      if (fl > 0) {
        const double* si = &M.dof_solimp[5 * i];
        const double* sr = &M.dof_solref[2 * i];
        const double imp = std::min(std::max(si[0], 0.0001), 0.9999), dmax = std::min(std::max(si[1], 0.0001), 0.9999);
        const double R = std::max(1e-15, (1 - imp) / imp * M.dof_invweight0[i]);
        double B;
        if (sr[0] > 0) { const double tc = std::max(sr[0], 2 * M.timestep); B = 2 / std::max(1e-15, dmax * tc); }
        else B = -sr[1] / std::max(1e-15, dmax);
        o[0] = R * fl; o[1] = 1 / R; o[2] = B;
      }
    }
  }
  M.body_dofmask.assign(nb, 0);
  for (int b = 1; b < nb; ++b) {
    unsigned mask = (unsigned)M.body_dofmask[M.body_parent[b]];
    for (int i = 0; i < M.body_dofnum[b]; ++i) mask |= 1u << (M.body_dofadr[b] + i);
    M.body_dofmask[b] = (int)mask;
  }
  // ---- elimination forests of the sparse (level-scheduled L^T D L) solver, dex_sparse.cuh.  Structure 0 = the dof
  //      forest itself (the pattern of M: contact-free envs, implicit-damping Euler of every env).  Structure 1 =
  //      "chained": the trees without branching (a free object's six dofs) strung into one chain that becomes the
  //      common ancestor of every branching tree; then H = M + J^T W J keeps the no-fill pattern "a dof couples to its
  //      ancestors only" for every contact whose dofs lie on one root-to-leaf path (object - finger, object - palm,
  //      finger - palm; not finger - finger).  TRUNK = the dofs above a branching point (wrist; object + wrist): a chain
  //      of up to 8 at depths 0 .. nt-1; every other dof has at most one child, so each lane pulls from at most one row
  //      per step, and the trunk takes the rows below it in one deferred (Schur) update.  Tables: DEX_SPL_* per lane.
  {
    M.sp_ok = 1;
    M.body_augmask.assign(nb, 0);
    std::vector<int> nchild(nv, 0), branching(M.ntree, 0);
    for (int i = 0; i < nv; ++i) if (M.dof_parent[i] >= 0) nchild[M.dof_parent[i]]++;
    for (int i = 0; i < nv; ++i) if (nchild[i] >= 2) branching[M.dof_tree[i]] = 1;
    std::vector<int> par[2];
    par[0] = M.dof_parent;
    par[1] = M.dof_parent;
    int tail = -1;   // deepest dof of the chain so far
    for (int i = 0; i < nv; ++i) if (!branching[M.dof_tree[i]]) { if (M.dof_parent[i] < 0) par[1][i] = tail; if (nchild[i] == 0) tail = i; }
    for (int i = 0; i < nv; ++i) if (branching[M.dof_tree[i]] && M.dof_parent[i] < 0) par[1][i] = tail;
    M.sp_tab.assign(2 * DEX_SPL_WORDS * 32, 0);
    for (int s = 0; s < 2 && M.sp_ok; ++s) {
      const std::vector<int>& P = par[s];
      DexSpConst& C = M.sp[s];
      std::memset(&C, 0, sizeof(C));
      std::vector<int> depth(32, 0), step(32, 255), nch(nv, 0), trunk(nv, 0);
      int maxd = 0;
      for (int i = 0; i < nv; ++i) {
        int d = 0, guard = 0;
        for (int j = P[i]; j >= 0 && guard < 64; j = P[j], ++guard) ++d;
        if (guard >= 64) { M.sp_ok = 0; break; }
        depth[i] = d; maxd = std::max(maxd, d);
        if (P[i] >= 0) nch[P[i]]++;
      }
      if (maxd > DEX_SP_SLOTS || maxd + 1 > 16) M.sp_ok = 0;
      if (!M.sp_ok) break;
      for (int i = 0; i < nv; ++i) if (nch[i] >= 2) for (int j = i; j >= 0; j = P[j]) trunk[j] = 1;
      int nt = 0;
      for (int i = 0; i < nv; ++i) if (trunk[i]) {
        nt++;
        int tc = 0;
        for (int k = 0; k < nv; ++k) if (P[k] == i && trunk[k]) tc++;
        if (tc > 1) M.sp_ok = 0;    // two branching points side by side: not a chain
      }
      if (nt > 8) M.sp_ok = 0;
      for (int i = 0; i < nv && M.sp_ok; ++i) if (trunk[i]) { if (depth[i] >= nt) M.sp_ok = 0; else C.trunk[depth[i]] = (unsigned char)i; }
      auto below_trunk = [&](int i) { for (int j = P[i]; j >= 0; j = P[j]) if (trunk[j]) return true; return false; };
      int ndef = 0;
      for (int i = 0; i < nv && M.sp_ok; ++i) if (!trunk[i] && below_trunk(i)) { if (ndef >= 24) { M.sp_ok = 0; break; } C.def[ndef++] = (unsigned char)i; }
      if (!M.sp_ok) break;
      C.nstep = maxd + 1; C.nslot = maxd; C.nt = nt; C.ndef = ndef;
      C.first_trunk_step = nt > 0 ? maxd - (nt - 1) : maxd + 1;
      for (int i = 0; i < nv; ++i) step[i] = maxd - depth[i];   // deepest dofs first
      int* T = &M.sp_tab[(size_t)s * DEX_SPL_WORDS * 32];
      for (int i = 0; i < 32; ++i) {
        T[DEX_SPL_DEPTH * 32 + i] = depth[i]; T[DEX_SPL_STEP * 32 + i] = step[i];
        unsigned char anc[16], d0[16];
        for (int d = 0; d < 16; ++d) { anc[d] = (unsigned char)i; d0[d] = 0xff; }
        if (i < nv) {
          int d = depth[i];
          for (int j = P[i]; j >= 0; j = P[j]) anc[--d] = (unsigned char)j;
          // the row a lane pulls at each step: its only descendant eliminated then (trunk lanes: trunk descendants only)
          for (int k = 0; k < nv; ++k) {
            bool isanc = false;
            for (int j = P[k]; j >= 0; j = P[j]) if (j == i) isanc = true;
            if (!isanc || (trunk[i] && !trunk[k])) continue;
            if (d0[step[k]] != 0xff) M.sp_ok = 0;   // cannot happen: non-trunk dofs have one child at most
            d0[step[k]] = (unsigned char)k;
          }
        }
        for (int w = 0; w < 4; ++w) {
          unsigned a = 0, b = 0;
          for (int q = 0; q < 4; ++q) { a |= (unsigned)anc[4 * w + q] << (8 * q); b |= (unsigned)d0[4 * w + q] << (8 * q); }
          T[(DEX_SPL_ANC + w) * 32 + i] = (int)a; T[(DEX_SPL_DESC + w) * 32 + i] = (int)b;
        }
        // Schur entries (t, j), 0 <= j <= t < nt, dealt to the lanes in order t (t + 1) / 2 + j; entries 32.. are second entries of lanes 0..
        unsigned sch = 0xffffu;
        int q = 0;
        for (int t = 0; t < nt; ++t) for (int j = 0; j <= t; ++j, ++q) {
          if (q == i) sch = (sch & ~0xffu) | (unsigned)t | ((unsigned)j << 4);
          if (q == i + 32) sch = (sch & ~0xff00u) | ((unsigned)t << 8) | ((unsigned)j << 12);
        }
        T[DEX_SPL_SCHUR * 32 + i] = (int)sch;
      }
    }
    if (!M.sp_ok) { std::memset(M.sp, 0, sizeof(M.sp)); }
    else {
      // dofs a contact of body b may touch and still fit structure 1: its own chain of ancestors there
      const std::vector<int>& P = par[1];
      for (int b = 1; b < nb; ++b) {
        unsigned mask = 0;
        int leaf = -1;   // deepest dof that moves b
        for (int a = b; a > 0 && leaf < 0; a = M.body_parent[a]) if (M.body_dofnum[a] > 0) leaf = M.body_dofadr[a] + M.body_dofnum[a] - 1;
        for (int j = leaf; j >= 0; j = P[j]) mask |= 1u << j;
        M.body_augmask[b] = (int)mask;
      }
    }
  }
  // reference body of each tree: the moving body whose origin (at qpos0) minimises the mean squared
  // distance to the tree's moving bodies -> small lever arms in FP32 (MuJoCo uses the subtree COM)
  {
    std::vector<double> xp(3 * nb, 0), xm(9 * nb, 0), xq(4 * nb, 0);
    xq[0] = 1; q2m(&xq[0], &xm[0]);
    for (int b = 1; b < nb; ++b) {
      int p = M.body_parent[b];
      const double* R = &xm[9 * p];
      const double* bp = &M.body_pos[3 * b];
      for (int k = 0; k < 3; ++k) xp[3 * b + k] = xp[3 * p + k] + R[3 * k] * bp[0] + R[3 * k + 1] * bp[1] + R[3 * k + 2] * bp[2];
      const double* a = &xq[4 * p];
      const double* q = &M.body_quat[4 * b];
      double r[4] = {a[0] * q[0] - a[1] * q[1] - a[2] * q[2] - a[3] * q[3], a[0] * q[1] + a[1] * q[0] + a[2] * q[3] - a[3] * q[2],
                     a[0] * q[2] - a[1] * q[3] + a[2] * q[0] + a[3] * q[1], a[0] * q[3] + a[1] * q[2] - a[2] * q[1] + a[3] * q[0]};
      if (M.body_jntnum[b] > 0 && M.jnt_type[M.body_jntadr[b]] == DEX_JNT_FREE) {
        int qa = M.jnt_qposadr[M.body_jntadr[b]];
        for (int k = 0; k < 3; ++k) xp[3 * b + k] = M.qpos0[qa + k];
        for (int k = 0; k < 4; ++k) r[k] = M.qpos0[qa + 3 + k];
      }
      std::memcpy(&xq[4 * b], r, sizeof(r));
      q2m(r, &xm[9 * b]);
    }
    M.tree_refbody.assign(M.ntree, 0);
    for (int t = 1; t < M.ntree; ++t) {
      double best = 1e300; int bestb = 0;
      for (int c = 1; c < nb; ++c) {
        if (M.body_tree[c] != t) continue;
        double s = 0; int n = 0;
        for (int b = 1; b < nb; ++b) if (M.body_tree[b] == t && M.body_weld[b] != 0) {
          double d2 = 0;
          for (int k = 0; k < 3; ++k) d2 += (xp[3 * b + k] - xp[3 * c + k]) * (xp[3 * b + k] - xp[3 * c + k]);
          s += d2; n++;
        }
        if (n == 0) { bestb = c; break; }
        if (s < best) { best = s; bestb = c; }
      }
      M.tree_refbody[t] = bestb;
    }
  }
  // ---- body-frame inertia matrices
  {
    std::vector<double> iq = F(DEXF_BODY_IQUAT), in = F(DEXF_BODY_INERTIA);
    M.body_inertia6.assign(6 * nb, 0);
    for (int b = 0; b < nb; ++b) {
      double R[9], A[9];
      q2m(&iq[4 * b], R);
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += R[3 * r + k] * in[3 * b + k] * R[3 * c + k];
        A[3 * r + c] = s;
      }
      double* o = &M.body_inertia6[6 * b];
      o[0] = A[0]; o[1] = A[4]; o[2] = A[8]; o[3] = A[1]; o[4] = A[2]; o[5] = A[5];
    }
  }
  // ---- packed constant records of the tree kernel K0 (vector loads instead of ~35 scalar loads per body)
  M.body_rec.assign(24 * nb, 0.0);
  for (int b = 0; b < nb; ++b) {
    double* r = &M.body_rec[24 * b];
    for (int k = 0; k < 3; ++k) { r[k] = M.body_pos[3 * b + k]; r[8 + k] = M.body_ipos[3 * b + k]; }
    r[3] = M.body_mass[b];
    for (int k = 0; k < 4; ++k) r[4 + k] = M.body_quat[4 * b + k];
    for (int k = 0; k < 6; ++k) { r[12 + k] = M.body_inertia6[6 * b + k]; r[18 + k] = M.body_xfrc[6 * b + k]; }
  }
  M.jnt_rec.assign(std::max(1, 8 * M.njnt), 0.0); M.jnt_irec.assign(std::max(1, 4 * M.njnt), 0);
  for (int j = 0; j < M.njnt; ++j) {
    for (int k = 0; k < 3; ++k) { M.jnt_rec[8 * j + k] = M.jnt_axis[3 * j + k]; M.jnt_rec[8 * j + 4 + k] = M.jnt_pos[3 * j + k]; }
    M.jnt_rec[8 * j + 3] = M.qpos0[M.jnt_qposadr[j]];
    M.jnt_irec[4 * j] = M.jnt_type[j]; M.jnt_irec[4 * j + 1] = M.jnt_qposadr[j]; M.jnt_irec[4 * j + 2] = M.jnt_dofadr[j];
  }
  // ---- geoms: body-frame rotation matrices and hull vertices
  {
    std::vector<double> gq = F(DEXF_GEOM_QUAT), v = F(DEXF_VERT);
    std::vector<int> va = I(DEXF_GEOM_VERTADR), vn = I(DEXF_GEOM_VERTNUM);
    M.geom_mat.assign(9 * M.ngeom, 0);
    M.geom_vertadr.assign(M.ngeom, -1);
    M.geom_vertnum = vn;
    M.vert.clear();
    for (int g = 0; g < M.ngeom; ++g) {
      double* G = &M.geom_mat[9 * g];
      q2m(&gq[4 * g], G);
      int t = M.geom_type[g];
      // spheres and capsules may be present (e.g. the Adroit hand) as long as no candidate pair names them: the pair
      // check below rejects every pair type without a narrowphase
      if (t != DEX_GEOM_PLANE && t != DEX_GEOM_BOX && t != DEX_GEOM_MESH && t != DEX_GEOM_SPHERE && t != DEX_GEOM_CAPSULE)
        return "unsupported geom type";
      if (t == DEX_GEOM_MESH) {
        M.geom_vertadr[g] = (int)(M.vert.size() / 4);
        for (int i = 0; i < vn[g]; ++i) {
          const double* p = &v[3 * (va[g] + i)];
          for (int k = 0; k < 3; ++k) M.vert.push_back(M.geom_pos[3 * g + k] + G[3 * k] * p[0] + G[3 * k + 1] * p[1] + G[3 * k + 2] * p[2]);
          M.vert.push_back(0.0);
        }
      }
    }
    // hull graph for the hill-climbing support function of the narrowphase: per MESH vertex one packed
    // word (first neighbour << 8 | neighbour count) into `adj` (neighbour ids local to the mesh), per geom
    // the mesh-vertex base of its hull and a cube-map table of start vertices
    {
      std::vector<int> aadr = I(DEXF_VERT_ADJADR), anum = I(DEXF_VERT_ADJNUM);
      M.adj = I(DEXF_ADJ);
      if (M.adj.empty()) M.adj.assign(1, 0);
      M.vadj.assign(std::max<size_t>(aadr.size(), 1), 0);
      for (size_t i = 0; i < aadr.size(); ++i) {
        if (anum[i] > 255 || aadr[i] >= (1 << 23)) return "hull graph exceeds the packed adjacency format";
        M.vadj[i] = (aadr[i] << 8) | anum[i];
      }
      M.geom_adjbase.assign(M.ngeom, 0);
      // start-vertex cube map: 6 faces x 8 x 8 cells, the exhaustive support vertex of each cell-centre direction,
      // two 16-bit vertex ids per word
      M.geom_cubemap.assign(DEX_CUBEMAP_WORDS * M.ngeom, 0);
      for (int g = 0; g < M.ngeom; ++g) {
        if (M.geom_type[g] != DEX_GEOM_MESH) continue;
        if (vn[g] >= 65536) return "hull too large for the 16-bit start-vertex table";
        M.geom_adjbase[g] = va[g];
        for (int c = 0; c < 384; ++c) {
          const int face = c >> 6, iu = (c >> 3) & 7, iv = c & 7;
          const double u = (iu + 0.5) / 4.0 - 1.0, v = (iv + 0.5) / 4.0 - 1.0, sgn = (face & 1) ? -1.0 : 1.0;
          double d[3];
          const int ax = face >> 1;
          d[ax] = sgn; d[(ax + 1) % 3] = u; d[(ax + 2) % 3] = v;
          int best = 0; double bv = -1e300;
          for (int i = 0; i < vn[g]; ++i) {
            const double* p = &M.vert[4 * (M.geom_vertadr[g] + i)];
            const double t = d[0] * p[0] + d[1] * p[1] + d[2] * p[2];
            if (t > bv) { bv = t; best = i; }
          }
          M.geom_cubemap[DEX_CUBEMAP_WORDS * g + (c >> 1)] |= best << (16 * (c & 1));
        }
        for (int i = 0; i < vn[g]; ++i) {
          if (va[g] + i >= (int)aadr.size()) return "hull graph missing for a mesh geom";
          for (int k = 0; k < anum[va[g] + i]; ++k) {
            const int j = M.adj[aadr[va[g] + i] + k];
            if (j < 0 || j >= vn[g]) return "hull graph neighbour out of range";
          }
        }
      }
    }
    // body-frame AABB (centre, half extents) of every geom: tightens the broadphase (conservative)
    M.geom_aabb.assign(6 * M.ngeom, 0);
    for (int g = 0; g < M.ngeom; ++g) {
      double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
      auto grow = [&](const double* p) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } };
      if (M.geom_type[g] == DEX_GEOM_MESH) {
        for (int i = 0; i < vn[g]; ++i) grow(&M.vert[4 * (M.geom_vertadr[g] + i)]);
      } else if (M.geom_type[g] == DEX_GEOM_BOX) {
        const double* G = &M.geom_mat[9 * g];
        for (int c = 0; c < 8; ++c) {
          double l[3] = {(c & 1 ? 1 : -1) * M.geom_size[3 * g], (c & 2 ? 1 : -1) * M.geom_size[3 * g + 1], (c & 4 ? 1 : -1) * M.geom_size[3 * g + 2]}, p[3];
          for (int k = 0; k < 3; ++k) p[k] = M.geom_pos[3 * g + k] + G[3 * k] * l[0] + G[3 * k + 1] * l[1] + G[3 * k + 2] * l[2];
          grow(p);
        }
      } else if (M.geom_type[g] == DEX_GEOM_CAPSULE || M.geom_type[g] == DEX_GEOM_SPHERE) {
        const double* G = &M.geom_mat[9 * g];
        const double r = M.geom_size[3 * g], hl = M.geom_type[g] == DEX_GEOM_CAPSULE ? M.geom_size[3 * g + 1] : 0.0;
        for (int sgn = -1; sgn <= 1; sgn += 2) for (int c = 0; c < 8; ++c) {   // boxes around the two end spheres
          double p[3];
          for (int k = 0; k < 3; ++k) p[k] = M.geom_pos[3 * g + k] + sgn * hl * G[3 * k + 2] + ((c >> k) & 1 ? r : -r);
          grow(p);
        }
      } else { for (int k = 0; k < 3; ++k) { lo[k] = -1e30; hi[k] = 1e30; } }
      for (int k = 0; k < 3; ++k) { M.geom_aabb[6 * g + k] = 0.5 * (lo[k] + hi[k]); M.geom_aabb[6 * g + 3 + k] = 0.5 * (hi[k] - lo[k]); }
    }
    // packed per-geom broadphase record (DEX_GPACK = 12): centre(3) rbound | aabb centre(3) half(3) | pad(2)
    M.geom_pack.assign(12 * M.ngeom, 0);
    for (int g = 0; g < M.ngeom; ++g) {
      double* pk = &M.geom_pack[12 * g];
      for (int k = 0; k < 3; ++k) { pk[k] = M.geom_pos[3 * g + k]; pk[4 + k] = M.geom_aabb[6 * g + k]; pk[7 + k] = M.geom_aabb[6 * g + 3 + k]; }
      pk[3] = M.geom_rbound[g];
    }
    if (M.vert.empty()) M.vert.assign(4, 0.0);
  }
  // ---- body bounding spheres (body frame); radius < 0 marks a body that owns a plane
  M.body_bsphere.assign(4 * nb, 0);
  for (int b = 0; b < nb; ++b) {
    double c[3] = {0, 0, 0}; int n = 0; bool plane = false;
    for (int g = 0; g < M.ngeom; ++g) if (M.geom_body[g] == b) {
      if (M.geom_type[g] == DEX_GEOM_PLANE) plane = true;
      for (int k = 0; k < 3; ++k) c[k] += M.geom_pos[3 * g + k];
      n++;
    }
    if (n) for (int k = 0; k < 3; ++k) c[k] /= n;
    double r = 0;
    for (int g = 0; g < M.ngeom; ++g) if (M.geom_body[g] == b) {
      double d2 = 0;
      for (int k = 0; k < 3; ++k) d2 += (M.geom_pos[3 * g + k] - c[k]) * (M.geom_pos[3 * g + k] - c[k]);
      r = std::max(r, std::sqrt(d2) + M.geom_rbound[g]);
    }
    for (int k = 0; k < 3; ++k) M.body_bsphere[4 * b + k] = c[k];
    M.body_bsphere[4 * b + 3] = plane ? (n == 1 ? -1.0 : 1e30) : r;
  }
  // ---- body-frame AABB over each body's collidable geoms, in the geom_pack record layout (only the
  //      aabb slots 4..9 are used) so the same OBB separating-axis routine serves bodies and geoms
  M.body_pack.assign(12 * nb, 0);
  for (int b = 0; b < nb; ++b) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    bool any = false;
    for (int g = 0; g < M.ngeom; ++g) if (M.geom_body[g] == b) {
      any = true;
      for (int k = 0; k < 3; ++k) {
        lo[k] = std::min(lo[k], M.geom_aabb[6 * g + k] - M.geom_aabb[6 * g + 3 + k]);
        hi[k] = std::max(hi[k], M.geom_aabb[6 * g + k] + M.geom_aabb[6 * g + 3 + k]);
      }
    }
    for (int k = 0; k < 3; ++k) {
      M.body_pack[12 * b + 4 + k] = any ? 0.5 * (lo[k] + hi[k]) : 0.0;
      M.body_pack[12 * b + 7 + k] = any ? std::min(0.5 * (hi[k] - lo[k]), 1e30) : 0.0;
    }
  }
  // ---- candidate pairs grouped by body pair (blob order is already grouped; verified here)
  M.pair_geom1 = I(DEXF_PAIR_GEOM1); M.pair_geom2 = I(DEXF_PAIR_GEOM2); M.pair_condim = I(DEXF_PAIR_CONDIM);
  M.pair_friction = F(DEXF_PAIR_FRICTION); M.pair_solref = F(DEXF_PAIR_SOLREF); M.pair_solimp = F(DEXF_PAIR_SOLIMP);
  M.pair_margin = F(DEXF_PAIR_MARGIN); M.pair_gap = F(DEXF_PAIR_GAP);
  {
    double maxmargin = 0;
    for (int p = 0; p < M.npair; ++p) {
      if (M.pair_condim[p] != 3 && M.pair_condim[p] != 1) return "only condim-1 and condim-3 contact pairs are supported yet";
      if (M.pair_condim[p] != 3) { M.arrow_ok = 0; M.sp_ok = 0; }   // the structured / sparse solver forms know pyramidal contacts only
      int t1 = M.geom_type[M.pair_geom1[p]], t2 = M.geom_type[M.pair_geom2[p]];
      // narrowphases: plane-box, plane-mesh, capsule-capsule (closed forms); any two of capsule / box / mesh through MPR
      // (geoms of a pair are ordered by type, mj_collideGeoms)
      auto convex = [](int t) { return t == DEX_GEOM_CAPSULE || t == DEX_GEOM_BOX || t == DEX_GEOM_MESH; };
      bool okp = (t1 == DEX_GEOM_PLANE && (t2 == DEX_GEOM_BOX || t2 == DEX_GEOM_MESH)) || (convex(t1) && convex(t2) && t1 <= t2 && !(t1 == DEX_GEOM_BOX && t2 == DEX_GEOM_BOX));
      if (!okp) return "unsupported geom pair type";
      maxmargin = std::max(maxmargin, M.pair_margin[p]);
      int a = M.geom_body[M.pair_geom1[p]], b = M.geom_body[M.pair_geom2[p]];
      if (M.bp_body1.empty() || M.bp_body1.back() != a || M.bp_body2.back() != b) {
        for (size_t q = 0; q < M.bp_body1.size(); ++q)
          if (M.bp_body1[q] == a && M.bp_body2[q] == b) return "candidate pairs are not grouped by body pair";
        M.bp_body1.push_back(a); M.bp_body2.push_back(b); M.bp_pairadr.push_back(p); M.bp_pairnum.push_back(0);
      }
      M.bp_pairnum.back()++;
    }
    M.pair_gg.assign(std::max(1, M.npair), 0);   // both geom ids in one word for the broadphase sweep
    for (int p = 0; p < M.npair; ++p) M.pair_gg[p] = M.pair_geom1[p] | (M.pair_geom2[p] << 16);
    M.nbodypair = (int)M.bp_body1.size();
    if (M.nbodypair > 384) return "more than 384 candidate body pairs (DEX_MAXCAND1)";
    if (M.ngeom > 256) return "more than 256 geoms (DEX_MAXGEOM)";
    M.max_margin = maxmargin;
    // fold the largest pair margin into the body spheres so the body-level cull stays conservative
    for (int b = 0; b < nb; ++b) if (M.body_bsphere[4 * b + 3] >= 0) M.body_bsphere[4 * b + 3] += 0.5 * maxmargin;
  }
  if (M.bp_body1.empty()) { M.bp_body1.assign(1, 0); M.bp_body2.assign(1, 0); M.bp_pairadr.assign(1, 0); M.bp_pairnum.assign(1, 0); }
  return "";
}

}  // namespace dex_host
#endif  // DEX_HOST_MODEL_H_
