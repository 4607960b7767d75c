// mjstep_oracle.cpp — FP64 CPU restatement of MuJoCo's mj_step for the feature set the
// dexterity suite's scenes use.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.  The product path
// (dexterity_b200/csrc) never links, imports or calls it.
//
// PARITY UNPINNED: the arithmetic of the reference's hot path lives in the un-vendored,
// unpinned third-party `mujoco` C engine (reference setup.py:25-33), which is not
// installable here (no wheel, no network) and the reference's own tests hold no mj_step
// golden vectors.  This file restates the published algorithm of the
// MuJoCo 2.2/2.3 engine stage by stage; each function names the upstream routine it
// follows.  It is validated by analytic invariants in tests/ (M vs its definition,
// RNE vs finite differences, energy conservation, KKT residual of the solver, MPR depth
// vs brute-force support-function search, resting-contact equilibrium).
//
// Reference call sites this oracle stands in for:
//   physics.step()      dexterity/manipulation/goals/fingertip_position.py:97, composer loop
//   physics.forward()   dexterity/models/hands/dexterous_hand.py:165
//   physics.data.contact  dexterity/utils/mujoco_collisions.py:106
//   mj_objectVelocity   dexterity/utils/mujoco_utils.py:27
//   mj_jacSite          dexterity/utils/mujoco_utils.py:55-60
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../include/dex_model_blob.h"

namespace {

typedef double R;
const R MINVAL = 1e-15;

// ----------------------------------------------------------------------------- small math
inline void cross(R* r, const R* a, const R* b) {
  R x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
inline R dot3(const R* a, const R* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline R norm3(const R* a) { return std::sqrt(dot3(a, a)); }
inline R normalize3(R* a) {
  R n = norm3(a);
  if (n < MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; } else { a[0] /= n; a[1] /= n; a[2] /= n; }
  return n;
}
inline void quat_mul(R* r, const R* a, const R* b) {
  R w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  R x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  R y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  R z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
inline void quat_normalize(R* q) {
  R n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; } else { for (int i = 0; i < 4; i++) q[i] /= n; }
}
inline void quat2mat(R* m, const R* q) {
  R w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
inline void rot_vec(R* r, const R* m, const R* v) {  // r = m v (row-major 3x3)
  R x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  R y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  R z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
inline void rotT_vec(R* r, const R* m, const R* v) {  // r = m^T v
  R x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  R y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  R z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
inline void mat_mul3(R* r, const R* a, const R* b) {
  R t[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
    t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  std::memcpy(r, t, sizeof(t));
}
inline void axisangle2quat(R* q, const R* axis, R ang) {
  R s = std::sin(ang * 0.5);
  q[0] = std::cos(ang * 0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}

// ----------------------------------------------------------------------------- model
struct Model {
  std::vector<uint8_t> blob;
  const int32_t* I(int f) const { return (const int32_t*)(blob.data() + ent(f).offset); }
  const R* F(int f) const { return (const R*)(blob.data() + ent(f).offset); }
  const DexBlobEntry& ent(int f) const { return ((const DexBlobEntry*)(blob.data() + sizeof(DexBlobHeader)))[f]; }
  int nbody, njnt, nq, nv, nu, ntendon, ngeom, npair, nsite, nconmax, iterations, mpr_iterations;
  R timestep, gravity[3], tolerance, mpr_tolerance, meaninertia;
};

struct Contact {
  R dist, pos[3], frame[9], includemargin, friction[5], solref[2], solimp[5];
  int dim, geom1, geom2;
};

enum { EFC_FRICTION = 0, EFC_LIMIT_JOINT = 1, EFC_LIMIT_TENDON = 2, EFC_CONTACT = 3 };

struct Data {
  const Model* m;
  std::vector<R> qpos, qvel, ctrl, qacc_warmstart;
  R time;
  std::vector<R> xpos, xquat, xmat, xipos, ximat, xanchor, xaxis, geom_xpos, geom_xmat, site_xpos, site_xmat;
  std::vector<R> subtree_com, cinert, cdof, cvel, cdof_dot;
  std::vector<R> M, qfrc_bias, qfrc_passive, qfrc_actuator, qfrc_applied, qfrc_smooth, qacc_smooth;
  std::vector<R> ten_length, ten_J, act_force;
  std::vector<Contact> contacts;
  int nefc;
  std::vector<R> efc_J, efc_pos, efc_margin, efc_D, efc_R, efc_aref, efc_floss, efc_force;
  std::vector<int> efc_type, efc_id;
  std::vector<R> qacc, qfrc_constraint;
  int solver_niter;
  int ncon_overflow;
  explicit Data(const Model* mm) : m(mm) {
    int nb = m->nbody, nv = m->nv;
    qpos.assign(m->F(DEXF_QPOS0), m->F(DEXF_QPOS0) + m->nq);
    qvel.assign(nv, 0); ctrl.assign(m->nu, 0); qacc_warmstart.assign(nv, 0); time = 0;
    xpos.assign(3 * nb, 0); xquat.assign(4 * nb, 0); xmat.assign(9 * nb, 0); xipos.assign(3 * nb, 0);
    ximat.assign(9 * nb, 0); xanchor.assign(3 * m->njnt, 0); xaxis.assign(3 * m->njnt, 0);
    geom_xpos.assign(3 * m->ngeom, 0); geom_xmat.assign(9 * m->ngeom, 0);
    site_xpos.assign(3 * m->nsite, 0); site_xmat.assign(9 * m->nsite, 0);
    subtree_com.assign(3 * nb, 0); cinert.assign(10 * nb, 0); cdof.assign(6 * nv, 0);
    cvel.assign(6 * nb, 0); cdof_dot.assign(6 * nv, 0);
    M.assign(nv * nv, 0); qfrc_bias.assign(nv, 0); qfrc_passive.assign(nv, 0); qfrc_actuator.assign(nv, 0);
    qfrc_applied.assign(nv, 0); qfrc_smooth.assign(nv, 0); qacc_smooth.assign(nv, 0);
    ten_length.assign(m->ntendon, 0); ten_J.assign(m->ntendon * nv, 0); act_force.assign(m->nu, 0);
    qacc.assign(nv, 0); qfrc_constraint.assign(nv, 0); nefc = 0; solver_niter = 0; ncon_overflow = 0;
  }
};

// ----------------------------------------------------------------------------- mj_kinematics
void kinematics(Data& d) {
  const Model& m = *d.m;
  const int32_t* parent = m.I(DEXF_BODY_PARENT);
  const R *bpos = m.F(DEXF_BODY_POS), *bquat = m.F(DEXF_BODY_QUAT);
  const R *bipos = m.F(DEXF_BODY_IPOS), *biquat = m.F(DEXF_BODY_IQUAT);
  const int32_t *jadr = m.I(DEXF_BODY_JNTADR), *jnum = m.I(DEXF_BODY_JNTNUM), *jtype = m.I(DEXF_JNT_TYPE);
  const int32_t* jqadr = m.I(DEXF_JNT_QPOSADR);
  const R *jpos = m.F(DEXF_JNT_POS), *jaxis = m.F(DEXF_JNT_AXIS), *qpos0 = m.F(DEXF_QPOS0);
  d.xpos[0] = d.xpos[1] = d.xpos[2] = 0;
  d.xquat[0] = 1; d.xquat[1] = d.xquat[2] = d.xquat[3] = 0;
  quat2mat(&d.xmat[0], &d.xquat[0]);
  for (int b = 1; b < m.nbody; b++) {
    int p = parent[b];
    R pos[3], quat[4], mat[9], t[3];
    rot_vec(t, &d.xmat[9 * p], bpos + 3 * b);
    for (int k = 0; k < 3; k++) pos[k] = d.xpos[3 * p + k] + t[k];
    quat_mul(quat, &d.xquat[4 * p], bquat + 4 * b);
    for (int j = jadr[b]; j >= 0 && j < jadr[b] + jnum[b]; j++) {
      int qa = jqadr[j];
      if (jtype[j] == 0) {  // free: pose is the qpos itself
        for (int k = 0; k < 3; k++) pos[k] = d.qpos[qa + k];
        for (int k = 0; k < 4; k++) quat[k] = d.qpos[qa + 3 + k];
        quat_normalize(quat);
        quat2mat(mat, quat);
        for (int k = 0; k < 3; k++) d.xanchor[3 * j + k] = pos[k];
        rot_vec(&d.xaxis[3 * j], mat, jaxis + 3 * j);
      } else {  // hinge
        quat2mat(mat, quat);
        rot_vec(t, mat, jpos + 3 * j);
        for (int k = 0; k < 3; k++) d.xanchor[3 * j + k] = pos[k] + t[k];
        rot_vec(&d.xaxis[3 * j], mat, jaxis + 3 * j);
        R ql[4], qn[4];
        axisangle2quat(ql, jaxis + 3 * j, d.qpos[qa] - qpos0[qa]);
        quat_mul(qn, quat, ql);
        std::memcpy(quat, qn, sizeof(qn));
        // correct for off-centre rotation
        quat2mat(mat, quat);
        rot_vec(t, mat, jpos + 3 * j);
        for (int k = 0; k < 3; k++) pos[k] = d.xanchor[3 * j + k] - t[k];
      }
    }
    quat_normalize(quat);
    quat2mat(mat, quat);
    std::memcpy(&d.xpos[3 * b], pos, sizeof(pos));
    std::memcpy(&d.xquat[4 * b], quat, sizeof(quat));
    std::memcpy(&d.xmat[9 * b], mat, sizeof(mat));
    rot_vec(t, mat, bipos + 3 * b);
    for (int k = 0; k < 3; k++) d.xipos[3 * b + k] = pos[k] + t[k];
    R iq[4];
    quat_mul(iq, quat, biquat + 4 * b);
    quat2mat(&d.ximat[9 * b], iq);
  }
  std::memcpy(&d.xipos[0], &d.xpos[0], 3 * sizeof(R));
  quat2mat(&d.ximat[0], &d.xquat[0]);
  const int32_t* gbody = m.I(DEXF_GEOM_BODY);
  const R *gpos = m.F(DEXF_GEOM_POS), *gquat = m.F(DEXF_GEOM_QUAT);
  for (int g = 0; g < m.ngeom; g++) {
    int b = gbody[g];
    R t[3], gm[9];
    rot_vec(t, &d.xmat[9 * b], gpos + 3 * g);
    for (int k = 0; k < 3; k++) d.geom_xpos[3 * g + k] = d.xpos[3 * b + k] + t[k];
    quat2mat(gm, gquat + 4 * g);
    mat_mul3(&d.geom_xmat[9 * g], &d.xmat[9 * b], gm);
  }
  const int32_t* sbody = m.I(DEXF_SITE_BODY);
  const R *spos = m.F(DEXF_SITE_POS), *squat = m.F(DEXF_SITE_QUAT);
  for (int s = 0; s < m.nsite; s++) {
    int b = sbody[s];
    R t[3], sm[9];
    rot_vec(t, &d.xmat[9 * b], spos + 3 * s);
    for (int k = 0; k < 3; k++) d.site_xpos[3 * s + k] = d.xpos[3 * b + k] + t[k];
    quat2mat(sm, squat + 4 * s);
    mat_mul3(&d.site_xmat[9 * s], &d.xmat[9 * b], sm);
  }
}

// ----------------------------------------------------------------------------- mj_comPos
// cinert layout (mju_inertCom): [Ixx Iyy Izz Ixy Ixz Iyz, m*dx m*dy m*dz, m], about subtree_com[root].
void com_pos(Data& d) {
  const Model& m = *d.m;
  const int32_t *parent = m.I(DEXF_BODY_PARENT), *root = m.I(DEXF_BODY_ROOT);
  const R *mass = m.F(DEXF_BODY_MASS), *inertia = m.F(DEXF_BODY_INERTIA);
  std::vector<R> sm(m.nbody, 0);
  for (int b = 0; b < m.nbody; b++) {
    for (int k = 0; k < 3; k++) d.subtree_com[3 * b + k] = d.xipos[3 * b + k] * mass[b];
    sm[b] = mass[b];
  }
  for (int b = m.nbody - 1; b > 0; b--) {
    int p = parent[b];
    for (int k = 0; k < 3; k++) d.subtree_com[3 * p + k] += d.subtree_com[3 * b + k];
    sm[p] += sm[b];
  }
  for (int b = 0; b < m.nbody; b++) {
    if (sm[b] < MINVAL) for (int k = 0; k < 3; k++) d.subtree_com[3 * b + k] = d.xipos[3 * b + k];
    else for (int k = 0; k < 3; k++) d.subtree_com[3 * b + k] /= sm[b];
  }
  for (int b = 1; b < m.nbody; b++) {
    R off[3];
    for (int k = 0; k < 3; k++) off[k] = d.xipos[3 * b + k] - d.subtree_com[3 * root[b] + k];
    const R* X = &d.ximat[9 * b];
    R I[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
      R s = 0;
      for (int k = 0; k < 3; k++) s += X[3 * i + k] * inertia[3 * b + k] * X[3 * j + k];
      I[3 * i + j] = s;
    }
    R mm = mass[b], o2 = dot3(off, off);
    R* c = &d.cinert[10 * b];
    c[0] = I[0] + mm * (o2 - off[0] * off[0]);
    c[1] = I[4] + mm * (o2 - off[1] * off[1]);
    c[2] = I[8] + mm * (o2 - off[2] * off[2]);
    c[3] = I[1] - mm * off[0] * off[1];
    c[4] = I[2] - mm * off[0] * off[2];
    c[5] = I[5] - mm * off[1] * off[2];
    c[6] = mm * off[0]; c[7] = mm * off[1]; c[8] = mm * off[2]; c[9] = mm;
  }
  // cdof: [angular(3); linear at subtree_com[root](3)] per dof
  const int32_t *jtype = m.I(DEXF_JNT_TYPE), *jbody = m.I(DEXF_JNT_BODY), *jdof = m.I(DEXF_JNT_DOFADR);
  for (int j = 0; j < m.njnt; j++) {
    int b = jbody[j], da = jdof[j];
    R off[3];
    for (int k = 0; k < 3; k++) off[k] = d.subtree_com[3 * root[b] + k] - d.xanchor[3 * j + k];
    if (jtype[j] == 0) {
      for (int k = 0; k < 3; k++) {
        R* c = &d.cdof[6 * (da + k)];
        for (int i = 0; i < 6; i++) c[i] = 0;
        c[3 + k] = 1;
      }
      for (int k = 0; k < 3; k++) {
        R ax[3] = {d.xmat[9 * b + k], d.xmat[9 * b + 3 + k], d.xmat[9 * b + 6 + k]};
        R* c = &d.cdof[6 * (da + 3 + k)];
        c[0] = ax[0]; c[1] = ax[1]; c[2] = ax[2];
        cross(c + 3, ax, off);
      }
    } else {
      R* c = &d.cdof[6 * da];
      for (int k = 0; k < 3; k++) c[k] = d.xaxis[3 * j + k];
      cross(c + 3, &d.xaxis[3 * j], off);
    }
  }
}

inline void mul_inert_vec(R* r, const R* c, const R* v) {  // mju_mulInertVec
  r[0] = c[0] * v[0] + c[3] * v[1] + c[4] * v[2] - c[8] * v[4] + c[7] * v[5];
  r[1] = c[3] * v[0] + c[1] * v[1] + c[5] * v[2] + c[8] * v[3] - c[6] * v[5];
  r[2] = c[4] * v[0] + c[5] * v[1] + c[2] * v[2] - c[7] * v[3] + c[6] * v[4];
  r[3] = c[8] * v[1] - c[7] * v[2] + c[9] * v[3];
  r[4] = c[6] * v[2] - c[8] * v[0] + c[9] * v[4];
  r[5] = c[7] * v[0] - c[6] * v[1] + c[9] * v[5];
}
inline void cross_motion(R* r, const R* v, const R* s) {  // mju_crossMotion: v x s
  R t[6];
  cross(t, v, s);
  R a[3], b[3];
  cross(a, v, s + 3);
  cross(b, v + 3, s);
  t[3] = a[0] + b[0]; t[4] = a[1] + b[1]; t[5] = a[2] + b[2];
  std::memcpy(r, t, sizeof(t));
}
inline void cross_force(R* r, const R* v, const R* f) {  // mju_crossForce: v x* f
  R a[3], b[3], c[3];
  cross(a, v, f);
  cross(b, v + 3, f + 3);
  cross(c, v, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}

// ----------------------------------------------------------------------------- mj_crb (dense M)
void crb(Data& d) {
  const Model& m = *d.m;
  const int32_t *parent = m.I(DEXF_BODY_PARENT), *dbody = m.I(DEXF_DOF_BODY), *dpar = m.I(DEXF_DOF_PARENT);
  const R* arm = m.F(DEXF_DOF_ARMATURE);
  int nv = m.nv;
  std::vector<R> c(d.cinert);
  for (int b = m.nbody - 1; b > 0; b--)
    if (parent[b] > 0) for (int k = 0; k < 10; k++) c[10 * parent[b] + k] += c[10 * b + k];
  std::fill(d.M.begin(), d.M.end(), 0.0);
  for (int i = 0; i < nv; i++) {
    R buf[6];
    mul_inert_vec(buf, &c[10 * dbody[i]], &d.cdof[6 * i]);
    for (int j = i; j >= 0; j = dpar[j]) {
      R s = 0;
      for (int k = 0; k < 6; k++) s += d.cdof[6 * j + k] * buf[k];
      d.M[i * nv + j] = d.M[j * nv + i] = s;
    }
    d.M[i * nv + i] += arm[i];
  }
}

// ----------------------------------------------------------------------------- mj_tendon / mj_transmission
void tendon(Data& d) {
  const Model& m = *d.m;
  const int32_t *tadr = m.I(DEXF_TEN_ADR), *tnum = m.I(DEXF_TEN_NUM), *wj = m.I(DEXF_WRAP_JNT);
  const int32_t *jq = m.I(DEXF_JNT_QPOSADR), *jd = m.I(DEXF_JNT_DOFADR);
  const R* wc = m.F(DEXF_WRAP_COEF);
  std::fill(d.ten_J.begin(), d.ten_J.end(), 0.0);
  for (int t = 0; t < m.ntendon; t++) {
    R L = 0;
    for (int w = tadr[t]; w < tadr[t] + tnum[t]; w++) {
      L += wc[w] * d.qpos[jq[wj[w]]];
      d.ten_J[t * m.nv + jd[wj[w]]] = wc[w];
    }
    d.ten_length[t] = L;
  }
}

// ----------------------------------------------------------------------------- mj_comVel
void com_vel(Data& d) {
  const Model& m = *d.m;
  const int32_t *parent = m.I(DEXF_BODY_PARENT), *jadr = m.I(DEXF_BODY_JNTADR), *jnum = m.I(DEXF_BODY_JNTNUM);
  const int32_t *jtype = m.I(DEXF_JNT_TYPE), *jdof = m.I(DEXF_JNT_DOFADR);
  for (int k = 0; k < 6; k++) d.cvel[k] = 0;
  for (int b = 1; b < m.nbody; b++) {
    R v[6];
    std::memcpy(v, &d.cvel[6 * parent[b]], sizeof(v));
    for (int j = jadr[b]; j >= 0 && j < jadr[b] + jnum[b]; j++) {
      int da = jdof[j];
      if (jtype[j] == 0) {
        for (int k = 0; k < 18; k++) d.cdof_dot[6 * da + k] = 0;
        for (int i = 0; i < 3; i++) for (int k = 0; k < 6; k++) v[k] += d.cdof[6 * (da + i) + k] * d.qvel[da + i];
        for (int i = 3; i < 6; i++) cross_motion(&d.cdof_dot[6 * (da + i)], v, &d.cdof[6 * (da + i)]);
        for (int i = 3; i < 6; i++) for (int k = 0; k < 6; k++) v[k] += d.cdof[6 * (da + i) + k] * d.qvel[da + i];
      } else {
        cross_motion(&d.cdof_dot[6 * da], v, &d.cdof[6 * da]);
        for (int k = 0; k < 6; k++) v[k] += d.cdof[6 * da + k] * d.qvel[da];
      }
    }
    std::memcpy(&d.cvel[6 * b], v, sizeof(v));
  }
}

// ----------------------------------------------------------------------------- mj_rne (flg_acc = 0)
void rne_bias(Data& d) {
  const Model& m = *d.m;
  const int32_t *parent = m.I(DEXF_BODY_PARENT), *dadr = m.I(DEXF_BODY_DOFADR), *dnum = m.I(DEXF_BODY_DOFNUM);
  const int32_t* dbody = m.I(DEXF_DOF_BODY);
  std::vector<R> cacc(6 * m.nbody, 0), cfrc(6 * m.nbody, 0);
  cacc[3] = -m.gravity[0]; cacc[4] = -m.gravity[1]; cacc[5] = -m.gravity[2];
  for (int b = 1; b < m.nbody; b++) {
    R a[6];
    std::memcpy(a, &cacc[6 * parent[b]], sizeof(a));
    for (int i = dadr[b]; i >= 0 && i < dadr[b] + dnum[b]; i++)
      for (int k = 0; k < 6; k++) a[k] += d.cdof_dot[6 * i + k] * d.qvel[i];
    std::memcpy(&cacc[6 * b], a, sizeof(a));
    R t1[6], t2[6], t3[6];
    mul_inert_vec(t1, &d.cinert[10 * b], a);
    mul_inert_vec(t2, &d.cinert[10 * b], &d.cvel[6 * b]);
    cross_force(t3, &d.cvel[6 * b], t2);
    for (int k = 0; k < 6; k++) cfrc[6 * b + k] = t1[k] + t3[k];
  }
  for (int b = m.nbody - 1; b > 0; b--)
    if (parent[b] > 0) for (int k = 0; k < 6; k++) cfrc[6 * parent[b] + k] += cfrc[6 * b + k];
  for (int i = 0; i < m.nv; i++) {
    R s = 0;
    for (int k = 0; k < 6; k++) s += d.cdof[6 * i + k] * cfrc[6 * dbody[i] + k];
    d.qfrc_bias[i] = s;
  }
}

// ----------------------------------------------------------------------------- mj_jac
// jacp/jacr: 3 x nv row-major. point attached to body.
void jac(const Data& d, R* jacp, R* jacr, const R* point, int body) {
  const Model& m = *d.m;
  const int32_t *parent = m.I(DEXF_BODY_PARENT), *root = m.I(DEXF_BODY_ROOT);
  const int32_t *dadr = m.I(DEXF_BODY_DOFADR), *dnum = m.I(DEXF_BODY_DOFNUM), *dpar = m.I(DEXF_DOF_PARENT);
  int nv = m.nv;
  if (jacp) std::fill(jacp, jacp + 3 * nv, 0.0);
  if (jacr) std::fill(jacr, jacr + 3 * nv, 0.0);
  while (body > 0 && dnum[body] == 0) body = parent[body];
  if (body <= 0) return;
  R off[3];
  for (int k = 0; k < 3; k++) off[k] = point[k] - d.subtree_com[3 * root[body] + k];
  for (int i = dadr[body] + dnum[body] - 1; i >= 0; i = dpar[i]) {
    const R* c = &d.cdof[6 * i];
    if (jacr) for (int k = 0; k < 3; k++) jacr[k * nv + i] = c[k];
    if (jacp) {
      R t[3];
      cross(t, c, off);
      for (int k = 0; k < 3; k++) jacp[k * nv + i] = c[3 + k] + t[k];
    }
  }
}

// ----------------------------------------------------------------------------- passive, actuation, acceleration
void passive(Data& d) {
  const R* damp = d.m->F(DEXF_DOF_DAMPING);
  for (int i = 0; i < d.m->nv; i++) d.qfrc_passive[i] = -damp[i] * d.qvel[i];
}

void fwd_actuation(Data& d) {
  const Model& m = *d.m;
  const int32_t *tt = m.I(DEXF_ACT_TRNTYPE), *ti = m.I(DEXF_ACT_TRNID);
  const int32_t *cl = m.I(DEXF_ACT_CTRLLIMITED), *fl = m.I(DEXF_ACT_FORCELIMITED);
  const int32_t *jq = m.I(DEXF_JNT_QPOSADR), *jd = m.I(DEXF_JNT_DOFADR);
  const R *gear = m.F(DEXF_ACT_GEAR), *gain = m.F(DEXF_ACT_GAINPRM), *bias = m.F(DEXF_ACT_BIASPRM);
  const R *cr = m.F(DEXF_ACT_CTRLRANGE), *fr = m.F(DEXF_ACT_FORCERANGE);
  int nv = m.nv;
  std::fill(d.qfrc_actuator.begin(), d.qfrc_actuator.end(), 0.0);
  for (int a = 0; a < m.nu; a++) {
    std::vector<R> moment(nv, 0);
    R length;
    if (tt[a] == 0) { length = gear[a] * d.qpos[jq[ti[a]]]; moment[jd[ti[a]]] = gear[a]; }
    else { length = gear[a] * d.ten_length[ti[a]]; for (int i = 0; i < nv; i++) moment[i] = gear[a] * d.ten_J[ti[a] * nv + i]; }
    R vel = 0;
    for (int i = 0; i < nv; i++) vel += moment[i] * d.qvel[i];
    R c = d.ctrl[a];
    if (cl[a]) c = std::min(std::max(c, cr[2 * a]), cr[2 * a + 1]);
    R f = gain[3 * a] * c + bias[3 * a] + bias[3 * a + 1] * length + bias[3 * a + 2] * vel;
    if (fl[a]) f = std::min(std::max(f, fr[2 * a]), fr[2 * a + 1]);
    d.act_force[a] = f;
    for (int i = 0; i < nv; i++) d.qfrc_actuator[i] += moment[i] * f;
  }
}

bool cholesky(std::vector<R>& A, int n) {  // in place lower, returns false if not PD
  for (int j = 0; j < n; j++) {
    R s = A[j * n + j];
    for (int k = 0; k < j; k++) s -= A[j * n + k] * A[j * n + k];
    if (s < MINVAL) return false;
    R l = std::sqrt(s);
    A[j * n + j] = l;
    for (int i = j + 1; i < n; i++) {
      R t = A[i * n + j];
      for (int k = 0; k < j; k++) t -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = t / l;
    }
  }
  return true;
}
void chol_solve(const std::vector<R>& L, int n, R* x) {
  for (int i = 0; i < n; i++) {
    R s = x[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k];
    x[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    R s = x[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}

void fwd_acceleration(Data& d) {
  const Model& m = *d.m;
  int nv = m.nv;
  const R* xfrc = m.F(DEXF_BODY_XFRC);
  // mj_xfrcAccumulate: force applied at the body com (xipos), torque directly
  std::fill(d.qfrc_applied.begin(), d.qfrc_applied.end(), 0.0);
  std::vector<R> jp(3 * nv), jr(3 * nv);
  for (int b = 1; b < m.nbody; b++) {
    const R* f = xfrc + 6 * b;
    bool any = false;
    for (int k = 0; k < 6; k++) any |= (f[k] != 0);
    if (!any) continue;
    jac(d, jp.data(), jr.data(), &d.xipos[3 * b], b);
    for (int i = 0; i < nv; i++)
      for (int k = 0; k < 3; k++) d.qfrc_applied[i] += jp[k * nv + i] * f[k] + jr[k * nv + i] * f[3 + k];
  }
  for (int i = 0; i < nv; i++)
    d.qfrc_smooth[i] = d.qfrc_passive[i] - d.qfrc_bias[i] + d.qfrc_applied[i] + d.qfrc_actuator[i];
  std::vector<R> L(d.M);
  cholesky(L, nv);
  d.qacc_smooth = d.qfrc_smooth;
  chol_solve(L, nv, d.qacc_smooth.data());
}

// ----------------------------------------------------------------------------- collision: support functions
struct Shape {
  const Data* d; int geom; R margin;
};
void support(const Shape& s, const R* dir, R* out) {  // mjccd_support (dir is unit length)
  const Model& m = *s.d->m;
  int g = s.geom;
  const R* pos = &s.d->geom_xpos[3 * g];
  const R* mat = &s.d->geom_xmat[9 * g];
  R ld[3], lr[3];
  rotT_vec(ld, mat, dir);
  int type = m.I(DEXF_GEOM_TYPE)[g];
  const R* size = m.F(DEXF_GEOM_SIZE) + 3 * g;
  if (type == 6) {
    for (int k = 0; k < 3; k++) lr[k] = (ld[k] > 0 ? size[k] : -size[k]);
  } else if (type == 7) {
    const R* v = m.F(DEXF_VERT) + 3 * m.I(DEXF_GEOM_VERTADR)[g];
    int n = m.I(DEXF_GEOM_VERTNUM)[g], best = 0;
    R bd = -DBL_MAX;
    for (int i = 0; i < n; i++) {
      R t = dot3(v + 3 * i, ld);
      if (t > bd) { bd = t; best = i; }
    }
    for (int k = 0; k < 3; k++) lr[k] = v[3 * best + k];
  } else if (type == 2) {
    for (int k = 0; k < 3; k++) lr[k] = ld[k] * size[0];
  } else if (type == 3) {
    for (int k = 0; k < 3; k++) lr[k] = ld[k] * size[0];
    lr[2] += (ld[2] > 0 ? size[1] : -size[1]);
  } else {
    lr[0] = lr[1] = lr[2] = 0;
  }
  rot_vec(out, mat, lr);
  for (int k = 0; k < 3; k++) out[k] += pos[k] + dir[k] * s.margin;
}

// ----------------------------------------------------------------------------- libccd MPR (ccdMPRPenetration)
const R CCD_EPS = DBL_EPSILON;
inline bool is_zero(R x) { return std::fabs(x) < CCD_EPS; }
inline bool ccd_eq(R a, R b) {
  R ab = std::fabs(a - b);
  if (ab < CCD_EPS) return true;
  R fa = std::fabs(a), fb = std::fabs(b);
  return fb > fa ? ab < CCD_EPS * fb : ab < CCD_EPS * fa;
}
struct SupPt { R v[3], v1[3], v2[3]; };
void mink_support(const Shape& a, const Shape& b, const R* dir, SupPt& p) {
  R nd[3] = {-dir[0], -dir[1], -dir[2]};
  support(a, dir, p.v1);
  support(b, nd, p.v2);
  for (int k = 0; k < 3; k++) p.v[k] = p.v1[k] - p.v2[k];
}
R point_seg_dist2(const R* P, const R* x0, const R* b, R* wit) {
  R dd[3], a[3];
  for (int k = 0; k < 3; k++) { dd[k] = b[k] - x0[k]; a[k] = x0[k] - P[k]; }
  R t = -dot3(a, dd) / dot3(dd, dd);
  R w[3];
  if (t < 0 || is_zero(t)) std::memcpy(w, x0, sizeof(w));
  else if (t > 1 || ccd_eq(t, 1)) std::memcpy(w, b, sizeof(w));
  else for (int k = 0; k < 3; k++) w[k] = x0[k] + t * dd[k];
  if (wit) std::memcpy(wit, w, sizeof(w));
  R df[3] = {w[0] - P[0], w[1] - P[1], w[2] - P[2]};
  return dot3(df, df);
}
R point_tri_dist2(const R* P, const R* x0, const R* B, const R* C, R* wit) {
  R d1[3], d2[3], a[3];
  for (int k = 0; k < 3; k++) { d1[k] = B[k] - x0[k]; d2[k] = C[k] - x0[k]; a[k] = x0[k] - P[k]; }
  R v = dot3(d1, d1), w = dot3(d2, d2), p = dot3(a, d1), q = dot3(a, d2), r = dot3(d1, d2);
  R s = (q * r - w * p) / (w * v - r * r);
  R t = (-s * r - q) / w;
  if ((is_zero(s) || s > 0) && (ccd_eq(s, 1) || s < 1) && (is_zero(t) || t > 0) && (ccd_eq(t, 1) || t < 1) &&
      (ccd_eq(t + s, 1) || t + s < 1)) {
    R wt[3];
    for (int k = 0; k < 3; k++) wt[k] = x0[k] + s * d1[k] + t * d2[k];
    std::memcpy(wit, wt, sizeof(wt));
    R df[3] = {wt[0] - P[0], wt[1] - P[1], wt[2] - P[2]};
    return dot3(df, df);
  }
  R w2[3];
  R dist = point_seg_dist2(P, x0, B, wit);
  R d2_ = point_seg_dist2(P, x0, C, w2);
  if (d2_ < dist) { dist = d2_; std::memcpy(wit, w2, sizeof(w2)); }
  d2_ = point_seg_dist2(P, B, C, w2);
  if (d2_ < dist) { dist = d2_; std::memcpy(wit, w2, sizeof(w2)); }
  return dist;
}
inline void portal_dir(const SupPt* s, R* dir) {
  R a[3], b[3];
  for (int k = 0; k < 3; k++) { a[k] = s[2].v[k] - s[1].v[k]; b[k] = s[3].v[k] - s[1].v[k]; }
  cross(dir, a, b);
  normalize3(dir);
}
inline bool portal_reach_tol(const SupPt* s, const SupPt& v4, const R* dir, R tol) {
  R dv1 = dot3(s[1].v, dir), dv2 = dot3(s[2].v, dir), dv3 = dot3(s[3].v, dir), dv4 = dot3(v4.v, dir);
  R m_ = std::min(dv4 - dv1, std::min(dv4 - dv2, dv4 - dv3));
  return ccd_eq(m_, tol) || m_ < tol;
}
inline void expand_portal(SupPt* s, const SupPt& v4) {
  R v4v0[3];
  cross(v4v0, v4.v, s[0].v);
  if (dot3(s[1].v, v4v0) > 0) {
    if (dot3(s[2].v, v4v0) > 0) s[1] = v4; else s[3] = v4;
  } else {
    if (dot3(s[3].v, v4v0) > 0) s[2] = v4; else s[1] = v4;
  }
}
void find_pos(const SupPt* s, R* pos) {
  R dir[3], b[4], t[3];
  portal_dir(s, dir);
  cross(t, s[1].v, s[2].v); b[0] = dot3(t, s[3].v);
  cross(t, s[3].v, s[2].v); b[1] = dot3(t, s[0].v);
  cross(t, s[0].v, s[1].v); b[2] = dot3(t, s[3].v);
  cross(t, s[2].v, s[1].v); b[3] = dot3(t, s[0].v);
  R sum = b[0] + b[1] + b[2] + b[3];
  if (is_zero(sum) || sum < 0) {
    b[0] = 0;
    cross(t, s[2].v, s[3].v); b[1] = dot3(t, dir);
    cross(t, s[3].v, s[1].v); b[2] = dot3(t, dir);
    cross(t, s[1].v, s[2].v); b[3] = dot3(t, dir);
    sum = b[1] + b[2] + b[3];
  }
  R inv = 1.0 / sum;
  for (int k = 0; k < 3; k++) {
    R p1 = 0, p2 = 0;
    for (int i = 0; i < 4; i++) { p1 += b[i] * s[i].v1[k]; p2 += b[i] * s[i].v2[k]; }
    pos[k] = 0.5 * (p1 + p2) * inv;
  }
}
// returns 0 on intersection (depth, dir, pos filled), -1 otherwise
int mpr_penetration(const Shape& o1, const Shape& o2, int max_iter, R tol, R* depth, R* pdir, R* pos) {
  const Data& d = *o1.d;
  SupPt s[4], v4;
  R dir[3], va[3], vb[3];
  // ---- discoverPortal
  for (int k = 0; k < 3; k++) {
    s[0].v1[k] = d.geom_xpos[3 * o1.geom + k];
    s[0].v2[k] = d.geom_xpos[3 * o2.geom + k];
    s[0].v[k] = s[0].v1[k] - s[0].v2[k];
  }
  if (is_zero(s[0].v[0]) && is_zero(s[0].v[1]) && is_zero(s[0].v[2])) s[0].v[0] += CCD_EPS * 10;
  for (int k = 0; k < 3; k++) dir[k] = -s[0].v[k];
  normalize3(dir);
  mink_support(o1, o2, dir, s[1]);
  R dt = dot3(s[1].v, dir);
  if (is_zero(dt) || dt < 0) return -1;
  cross(dir, s[0].v, s[1].v);
  int res = 0;
  if (is_zero(dot3(dir, dir))) {
    res = (is_zero(s[1].v[0]) && is_zero(s[1].v[1]) && is_zero(s[1].v[2])) ? 1 : 2;
  } else {
    normalize3(dir);
    mink_support(o1, o2, dir, s[2]);
    dt = dot3(s[2].v, dir);
    if (is_zero(dt) || dt < 0) return -1;
    for (int k = 0; k < 3; k++) { va[k] = s[1].v[k] - s[0].v[k]; vb[k] = s[2].v[k] - s[0].v[k]; }
    cross(dir, va, vb);
    normalize3(dir);
    if (dot3(dir, s[0].v) > 0) {
      std::swap(s[1], s[2]);
      for (int k = 0; k < 3; k++) dir[k] = -dir[k];
    }
    int size = 3, guard = 0;
    while (size < 4) {
      if (++guard > 1000) return -1;
      mink_support(o1, o2, dir, s[3]);
      dt = dot3(s[3].v, dir);
      if (is_zero(dt) || dt < 0) return -1;
      bool cont = false;
      cross(va, s[1].v, s[3].v);
      dt = dot3(va, s[0].v);
      if (dt < 0 && !is_zero(dt)) { s[2] = s[3]; cont = true; }
      if (!cont) {
        cross(va, s[3].v, s[2].v);
        dt = dot3(va, s[0].v);
        if (dt < 0 && !is_zero(dt)) { s[1] = s[3]; cont = true; }
      }
      if (cont) {
        for (int k = 0; k < 3; k++) { va[k] = s[1].v[k] - s[0].v[k]; vb[k] = s[2].v[k] - s[0].v[k]; }
        cross(dir, va, vb);
        normalize3(dir);
      } else {
        size = 4;
      }
    }
  }
  if (res == 1) {  // findPenetrTouch
    *depth = 0;
    pdir[0] = pdir[1] = pdir[2] = 0;
    for (int k = 0; k < 3; k++) pos[k] = 0.5 * (s[1].v1[k] + s[1].v2[k]);
    return 0;
  }
  if (res == 2) {  // findPenetrSegment
    for (int k = 0; k < 3; k++) pos[k] = 0.5 * (s[1].v1[k] + s[1].v2[k]);
    *depth = norm3(s[1].v);
    for (int k = 0; k < 3; k++) pdir[k] = s[1].v[k];
    normalize3(pdir);
    return 0;
  }
  // ---- refinePortal
  for (int guard = 0;; guard++) {
    if (guard > 1000) return -1;
    portal_dir(s, dir);
    dt = dot3(dir, s[1].v);
    if (is_zero(dt) || dt > 0) break;  // portal encapsules origin
    mink_support(o1, o2, dir, v4);
    dt = dot3(v4.v, dir);
    if (!(is_zero(dt) || dt > 0) || portal_reach_tol(s, v4, dir, tol)) return -1;
    expand_portal(s, v4);
  }
  // ---- findPenetr
  for (int it = 0;; it++) {
    portal_dir(s, dir);
    mink_support(o1, o2, dir, v4);
    if (portal_reach_tol(s, v4, dir, tol) || it > max_iter) {
      R origin[3] = {0, 0, 0};
      R d2 = point_tri_dist2(origin, s[1].v, s[2].v, s[3].v, pdir);
      *depth = std::sqrt(d2);
      if (is_zero(pdir[0]) && is_zero(pdir[1]) && is_zero(pdir[2])) std::memcpy(pdir, dir, 3 * sizeof(R));
      normalize3(pdir);
      find_pos(s, pos);
      return 0;
    }
    expand_portal(s, v4);
  }
}

// ----------------------------------------------------------------------------- narrowphase
int plane_box(const Data& d, int g1, int g2, R margin, Contact* con) {  // mjc_PlaneBox
  const Model& m = *d.m;
  const R *pos1 = &d.geom_xpos[3 * g1], *mat1 = &d.geom_xmat[9 * g1];
  const R *pos2 = &d.geom_xpos[3 * g2], *mat2 = &d.geom_xmat[9 * g2];
  const R* size2 = m.F(DEXF_GEOM_SIZE) + 3 * g2;
  R norm[3] = {mat1[2], mat1[5], mat1[8]}, dif[3];
  for (int k = 0; k < 3; k++) dif[k] = pos2[k] - pos1[k];
  R dist = dot3(dif, norm);
  int cnt = 0;
  for (int i = 0; i < 8; i++) {
    R vec[3] = {(i & 1 ? size2[0] : -size2[0]), (i & 2 ? size2[1] : -size2[1]), (i & 4 ? size2[2] : -size2[2])};
    R corner[3];
    rot_vec(corner, mat2, vec);
    R ldist = dot3(norm, corner);
    if (dist + ldist > margin || ldist > 0) continue;
    Contact& c = con[cnt];
    c.dist = dist + ldist;
    for (int k = 0; k < 3; k++) { c.frame[k] = norm[k]; c.frame[3 + k] = 0; }
    for (int k = 0; k < 3; k++) c.pos[k] = corner[k] + pos2[k] - norm[k] * c.dist * 0.5;
    if (++cnt >= 4) return 4;
  }
  return cnt;
}

int plane_convex(const Data& d, int g1, int g2, R margin, Contact* con) {  // mjc_PlaneConvex (mesh)
  const Model& m = *d.m;
  const R *pos1 = &d.geom_xpos[3 * g1], *mat1 = &d.geom_xmat[9 * g1];
  const R *pos2 = &d.geom_xpos[3 * g2], *mat2 = &d.geom_xmat[9 * g2];
  R norm[3] = {mat1[2], mat1[5], mat1[8]};
  R ld[3], nn[3] = {-norm[0], -norm[1], -norm[2]};
  rotT_vec(ld, mat2, nn);
  const R* v = m.F(DEXF_VERT) + 3 * m.I(DEXF_GEOM_VERTADR)[g2];
  int n = m.I(DEXF_GEOM_VERTNUM)[g2], best = 0;
  R bd = -DBL_MAX;
  for (int i = 0; i < n; i++) { R t = dot3(v + 3 * i, ld); if (t > bd) { bd = t; best = i; } }
  int cnt = 0;
  auto add = [&](int vi) {
    R w[3], dif[3];
    rot_vec(w, mat2, v + 3 * vi);
    for (int k = 0; k < 3; k++) { w[k] += pos2[k]; dif[k] = w[k] - pos1[k]; }
    R dist = dot3(dif, norm);
    if (dist > margin) return;
    Contact& c = con[cnt++];
    c.dist = dist;
    for (int k = 0; k < 3; k++) { c.frame[k] = norm[k]; c.frame[3 + k] = 0; c.pos[k] = w[k] - 0.5 * dist * norm[k]; }
  };
  add(best);
  if (cnt == 0) return 0;
  // up to 3 more: hull-graph neighbours of the support vertex within margin
  int va = m.I(DEXF_GEOM_VERTADR)[g2];
  const int32_t *aadr = m.I(DEXF_VERT_ADJADR), *anum = m.I(DEXF_VERT_ADJNUM), *adj = m.I(DEXF_ADJ);
  for (int k = 0; k < anum[va + best] && cnt < 4; k++) add(adj[aadr[va + best] + k]);
  return cnt;
}

int convex_convex(const Data& d, int g1, int g2, R margin, Contact* con) {  // mjc_Convex
  Shape a{&d, g1, 0.5 * margin}, b{&d, g2, 0.5 * margin};
  R depth, dir[3], pos[3];
  if (mpr_penetration(a, b, d.m->mpr_iterations, d.m->mpr_tolerance, &depth, dir, pos) != 0) return 0;
  if (dir[0] == 0 && dir[1] == 0 && dir[2] == 0) return 0;
  con->dist = margin - depth;
  for (int k = 0; k < 3; k++) { con->frame[k] = dir[k]; con->frame[3 + k] = 0; con->pos[k] = pos[k]; }
  return 1;
}

// mjraw_SphereSphere [upstream engine_collision_primitive.c]
int sphere_sphere(const R* p1, R r1, const R* p2, R r2, R margin, Contact* con) {
  R dif[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  R cdist = norm3(dif);
  if (cdist - r1 - r2 > margin) return 0;
  con->dist = cdist - r1 - r2;
  R n[3] = {1, 0, 0};
  if (cdist >= MINVAL) for (int k = 0; k < 3; k++) n[k] = dif[k] / cdist;
  for (int k = 0; k < 3; k++) { con->frame[k] = n[k]; con->frame[3 + k] = 0; con->pos[k] = p1[k] + n[k] * (r1 + 0.5 * con->dist); }
  return 1;
}
// mjc_CapsuleCapsule [upstream engine_collision_primitive.c]: closest points of the two axis segments (2 x 2 system in the
// segment parameters, clamped), then sphere-sphere; parallel axes: the end points of each segment against the other, two
// contacts at most.  Used by the Adroit hand's explicit <contact><pair>s (adroit_hand.xml:231-251).
int capsule_capsule(const Data& d, int g1, int g2, R margin, Contact* con) {
  const Model& m = *d.m;
  const R *pos1 = &d.geom_xpos[3 * g1], *pos2 = &d.geom_xpos[3 * g2], *mat1 = &d.geom_xmat[9 * g1], *mat2 = &d.geom_xmat[9 * g2];
  const R *size1 = m.F(DEXF_GEOM_SIZE) + 3 * g1, *size2 = m.F(DEXF_GEOM_SIZE) + 3 * g2;
  R ax1[3] = {mat1[2], mat1[5], mat1[8]}, ax2[3] = {mat2[2], mat2[5], mat2[8]};
  R dif[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
  R ma = dot3(ax1, ax1), mb = -dot3(ax1, ax2), mc = dot3(ax2, ax2);
  R u = -dot3(ax1, dif), v = dot3(ax2, dif);
  R det = ma * mc - mb * mb;
  auto clip = [](R x, R lo, R hi) { return std::min(std::max(x, lo), hi); };
  int n = 0;
  auto at = [&](R x1, R x2) {
    R v1[3], v2[3];
    for (int k = 0; k < 3; k++) { v1[k] = pos1[k] + ax1[k] * x1; v2[k] = pos2[k] + ax2[k] * x2; }
    if (n < 2) n += sphere_sphere(v1, size1[0], v2, size2[0], margin, con + n);
  };
  if (std::fabs(det) >= MINVAL) {
    R x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
    if (x1 > size1[1]) { x1 = size1[1]; x2 = (v - mb * size1[1]) / mc; }
    else if (x1 < -size1[1]) { x1 = -size1[1]; x2 = (v + mb * size1[1]) / mc; }
    if (x2 > size2[1]) { x2 = size2[1]; x1 = clip((u - mb * size2[1]) / ma, -size1[1], size1[1]); }
    else if (x2 < -size2[1]) { x2 = -size2[1]; x1 = clip((u + mb * size2[1]) / ma, -size1[1], size1[1]); }
    at(x1, x2);
  } else {
    at(size1[1], clip((v - mb * size1[1]) / mc, -size2[1], size2[1]));
    at(-size1[1], clip((v + mb * size1[1]) / mc, -size2[1], size2[1]));
    if (n < 2) at(clip((u - mb * size2[1]) / ma, -size1[1], size1[1]), size2[1]);
    if (n < 2) at(clip((u + mb * size2[1]) / ma, -size1[1], size1[1]), -size2[1]);
  }
  return n;
}

void make_frame(R* f) {  // mju_makeFrame
  normalize3(f);
  if (norm3(f + 3) < 0.5) {
    f[3] = f[4] = f[5] = 0;
    if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1;
  }
  R dp = dot3(f, f + 3);
  for (int k = 0; k < 3; k++) f[3 + k] -= dp * f[k];
  normalize3(f + 3);
  cross(f + 6, f, f + 3);
}

// ----------------------------------------------------------------------------- mj_collision
void collision(Data& d) {
  const Model& m = *d.m;
  const int32_t *g1s = m.I(DEXF_PAIR_GEOM1), *g2s = m.I(DEXF_PAIR_GEOM2), *gtype = m.I(DEXF_GEOM_TYPE);
  const R* rb = m.F(DEXF_GEOM_RBOUND);
  d.contacts.clear();
  d.ncon_overflow = 0;
  for (int p = 0; p < m.npair; p++) {
    int g1 = g1s[p], g2 = g2s[p];
    R margin = m.F(DEXF_PAIR_MARGIN)[p];
    // bounding-sphere filter (mj_collideGeoms / mj_filterSphere)
    if (gtype[g1] == 0) {
      const R* mat1 = &d.geom_xmat[9 * g1];
      R norm[3] = {mat1[2], mat1[5], mat1[8]}, dif[3];
      for (int k = 0; k < 3; k++) dif[k] = d.geom_xpos[3 * g2 + k] - d.geom_xpos[3 * g1 + k];
      if (dot3(dif, norm) > margin + rb[g2]) continue;
    } else {
      R dif[3];
      for (int k = 0; k < 3; k++) dif[k] = d.geom_xpos[3 * g2 + k] - d.geom_xpos[3 * g1 + k];
      R bound = margin + rb[g1] + rb[g2];
      if (dot3(dif, dif) > bound * bound) continue;
    }
    Contact con[4];
    int n = 0;
    if (gtype[g1] == 0 && gtype[g2] == 6) n = plane_box(d, g1, g2, margin, con);
    else if (gtype[g1] == 0 && gtype[g2] == 7) n = plane_convex(d, g1, g2, margin, con);
    else if (gtype[g1] == 3 && gtype[g2] == 3) n = capsule_capsule(d, g1, g2, margin, con);
    // capsule-box goes through MPR here AND in the kernels: upstream has a closed form (mjc_CapsuleBox, up to two contacts) that
    // this restatement does not reproduce; one of the Adroit hand's 19 pairs (C_palm0 - C_thdistal) is of this type
    else if ((gtype[g1] == 3 || gtype[g1] == 6 || gtype[g1] == 7) && (gtype[g2] == 6 || gtype[g2] == 7) && !(gtype[g1] == 6 && gtype[g2] == 6))
      n = convex_convex(d, g1, g2, margin, con);
    else continue;  // pair type not used by the compiled scenes
    for (int i = 0; i < n; i++) {
      Contact& c = con[i];
      c.geom1 = g1; c.geom2 = g2;
      c.dim = m.I(DEXF_PAIR_CONDIM)[p];
      c.includemargin = margin - m.F(DEXF_PAIR_GAP)[p];
      std::memcpy(c.friction, m.F(DEXF_PAIR_FRICTION) + 5 * p, 5 * sizeof(R));
      std::memcpy(c.solref, m.F(DEXF_PAIR_SOLREF) + 2 * p, 2 * sizeof(R));
      std::memcpy(c.solimp, m.F(DEXF_PAIR_SOLIMP) + 5 * p, 5 * sizeof(R));
      make_frame(c.frame);
      if ((int)d.contacts.size() >= m.nconmax) { d.ncon_overflow++; continue; }
      d.contacts.push_back(c);
    }
  }
}

// ----------------------------------------------------------------------------- mj_makeConstraint + mj_makeImpedance + mj_referenceConstraint
void get_imp(const R* solimp, R pos, R margin, R* imp) {  // getimpedance
  R dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  dmin = std::min(std::max(dmin, 0.0001), 0.9999);
  dmax = std::min(std::max(dmax, 0.0001), 0.9999);
  width = std::max(width, MINVAL);
  mid = std::min(std::max(mid, 0.0001), 0.9999);
  power = std::max(power, 1.0);
  R x = std::fabs((pos - margin) / width);
  if (x >= 1) { *imp = dmax; return; }
  if (x <= 0) { *imp = dmin; return; }
  R y;
  if (power == 1) y = x;
  else if (x <= mid) y = std::pow(x / mid, power) * mid;  // a*x^p with a = 1/mid^(p-1)
  else y = 1 - std::pow((1 - x) / (1 - mid), power) * (1 - mid);
  *imp = dmin + y * (dmax - dmin);
}

void make_constraint(Data& d) {
  const Model& m = *d.m;
  int nv = m.nv;
  d.efc_J.clear(); d.efc_pos.clear(); d.efc_margin.clear(); d.efc_type.clear(); d.efc_id.clear(); d.efc_floss.clear();
  auto add_row = [&](const std::vector<R>& J, R pos, R margin, R floss, int type, int id) {
    d.efc_J.insert(d.efc_J.end(), J.begin(), J.end());
    d.efc_pos.push_back(pos); d.efc_margin.push_back(margin); d.efc_floss.push_back(floss);
    d.efc_type.push_back(type); d.efc_id.push_back(id);
  };
  std::vector<R> row(nv);
  // dof friction loss
  const R* fl = m.F(DEXF_DOF_FRICTIONLOSS);
  for (int i = 0; i < nv; i++) if (fl[i] > 0) {
    std::fill(row.begin(), row.end(), 0.0);
    row[i] = 1;
    add_row(row, 0, 0, fl[i], EFC_FRICTION, i);
  }
  // joint limits (hinge/slide)
  const int32_t *jtype = m.I(DEXF_JNT_TYPE), *jlim = m.I(DEXF_JNT_LIMITED), *jq = m.I(DEXF_JNT_QPOSADR), *jd = m.I(DEXF_JNT_DOFADR);
  const R *jrange = m.F(DEXF_JNT_RANGE), *jmargin = m.F(DEXF_JNT_MARGIN);
  for (int j = 0; j < m.njnt; j++) if (jlim[j] && jtype[j] >= 2) {
    R value = d.qpos[jq[j]];
    for (int side = -1; side <= 1; side += 2) {
      R dist = side * (jrange[2 * j + (side + 1) / 2] - value);
      if (dist < jmargin[j]) {
        std::fill(row.begin(), row.end(), 0.0);
        row[jd[j]] = -(R)side;
        add_row(row, dist, jmargin[j], 0, EFC_LIMIT_JOINT, j);
      }
    }
  }
  // tendon limits
  const int32_t* tlim = m.I(DEXF_TEN_LIMITED);
  const R *trange = m.F(DEXF_TEN_RANGE), *tmargin = m.F(DEXF_TEN_MARGIN);
  for (int t = 0; t < m.ntendon; t++) if (tlim[t]) {
    R value = d.ten_length[t];
    for (int side = -1; side <= 1; side += 2) {
      R dist = side * (trange[2 * t + (side + 1) / 2] - value);
      if (dist < tmargin[t]) {
        for (int i = 0; i < nv; i++) row[i] = -(R)side * d.ten_J[t * nv + i];
        add_row(row, dist, tmargin[t], 0, EFC_LIMIT_TENDON, t);
      }
    }
  }
  // contacts (pyramidal cone; condim 1 → one frictionless row)
  const int32_t* gbody = m.I(DEXF_GEOM_BODY);
  std::vector<R> jp1(3 * nv), jp2(3 * nv), jd_(3 * nv), jc(3 * nv);
  for (size_t ci = 0; ci < d.contacts.size(); ci++) {
    const Contact& c = d.contacts[ci];
    jac(d, jp1.data(), nullptr, c.pos, gbody[c.geom1]);
    jac(d, jp2.data(), nullptr, c.pos, gbody[c.geom2]);
    for (int i = 0; i < 3 * nv; i++) jd_[i] = jp2[i] - jp1[i];
    for (int r = 0; r < 3; r++) for (int i = 0; i < nv; i++)
      jc[r * nv + i] = c.frame[3 * r] * jd_[i] + c.frame[3 * r + 1] * jd_[nv + i] + c.frame[3 * r + 2] * jd_[2 * nv + i];
    if (c.dim == 1) {
      for (int i = 0; i < nv; i++) row[i] = jc[i];
      add_row(row, c.dist, c.includemargin, 0, EFC_CONTACT, (int)ci);
    } else {
      for (int k = 1; k < c.dim && k < 3; k++) {
        for (int sgn = 1; sgn >= -1; sgn -= 2) {
          for (int i = 0; i < nv; i++) row[i] = jc[i] + sgn * c.friction[k - 1] * jc[k * nv + i];
          add_row(row, c.dist, c.includemargin, 0, EFC_CONTACT, (int)ci);
        }
      }
    }
  }
  int nefc = d.nefc = (int)d.efc_pos.size();
  d.efc_D.assign(nefc, 0); d.efc_R.assign(nefc, 0); d.efc_aref.assign(nefc, 0); d.efc_force.assign(nefc, 0);
  // diagApprox, impedance, R, K/B, aref
  const R *dinv = m.F(DEXF_DOF_INVWEIGHT0), *binv = m.F(DEXF_BODY_INVWEIGHT0), *tinv = m.F(DEXF_TEN_INVWEIGHT0);
  std::vector<R> KBIP(4 * nefc);
  for (int i = 0; i < nefc; i++) {
    int id = d.efc_id[i];
    R diag, sr[2], si[5];
    switch (d.efc_type[i]) {
      case EFC_FRICTION:
        diag = dinv[id];
        std::memcpy(sr, m.F(DEXF_DOF_SOLREF) + 2 * id, sizeof(sr));
        std::memcpy(si, m.F(DEXF_DOF_SOLIMP) + 5 * id, sizeof(si));
        break;
      case EFC_LIMIT_JOINT:
        diag = dinv[jd[id]];
        std::memcpy(sr, m.F(DEXF_JNT_SOLREF) + 2 * id, sizeof(sr));
        std::memcpy(si, m.F(DEXF_JNT_SOLIMP) + 5 * id, sizeof(si));
        break;
      case EFC_LIMIT_TENDON:
        diag = tinv[id];
        std::memcpy(sr, m.F(DEXF_TEN_SOLREF) + 2 * id, sizeof(sr));
        std::memcpy(si, m.F(DEXF_TEN_SOLIMP) + 5 * id, sizeof(si));
        break;
      default: {
        const Contact& c = d.contacts[id];
        int b1 = gbody[c.geom1], b2 = gbody[c.geom2];
        R tran = binv[2 * b1] + binv[2 * b2];
        diag = (c.dim == 1) ? tran : tran + c.friction[0] * c.friction[0] * tran;
        std::memcpy(sr, c.solref, sizeof(sr));
        std::memcpy(si, c.solimp, sizeof(si));
      }
    }
    if (sr[0] > 0) sr[0] = std::max(sr[0], 2 * m.timestep);  // refsafe
    R imp;
    get_imp(si, d.efc_pos[i], d.efc_margin[i], &imp);
    d.efc_R[i] = std::max(MINVAL, (1 - imp) / imp * diag);
    R dmax = std::min(std::max(si[1], 0.0001), 0.9999);
    R K, B;
    if (sr[0] > 0) { K = 1 / std::max(MINVAL, dmax * dmax * sr[0] * sr[0] * sr[1] * sr[1]); B = 2 / std::max(MINVAL, dmax * sr[0]); }
    else { K = -sr[0] / std::max(MINVAL, dmax * dmax); B = -sr[1] / std::max(MINVAL, dmax); }
    if (d.efc_type[i] == EFC_FRICTION) K = 0;
    KBIP[4 * i] = K; KBIP[4 * i + 1] = B; KBIP[4 * i + 2] = imp;
  }
  // pyramidal contacts: all rows of a contact share Rpy = 2 mu^2 R(first row)
  for (int i = 0; i < nefc;) {
    if (d.efc_type[i] == EFC_CONTACT && d.contacts[d.efc_id[i]].dim > 1) {
      const Contact& c = d.contacts[d.efc_id[i]];
      int nr = 2 * (std::min(c.dim, 3) - 1);
      R Rpy = 2 * c.friction[0] * c.friction[0] * d.efc_R[i];
      for (int k = 0; k < nr; k++) d.efc_R[i + k] = Rpy;
      i += nr;
    } else i++;
  }
  for (int i = 0; i < nefc; i++) {
    d.efc_D[i] = 1 / d.efc_R[i];
    R vel = 0;
    for (int k = 0; k < nv; k++) vel += d.efc_J[i * nv + k] * d.qvel[k];
    d.efc_aref[i] = -KBIP[4 * i + 1] * vel - KBIP[4 * i] * KBIP[4 * i + 2] * (d.efc_pos[i] - d.efc_margin[i]);
  }
}

// ----------------------------------------------------------------------------- mj_fwdConstraint: Newton (mj_solNewton)
struct SolverCtx {
  Data& d; int nv, nefc;
  std::vector<R> jar, Ma, grad, search, Mv, jv, active;
  R cost;
  explicit SolverCtx(Data& dd) : d(dd), nv(dd.m->nv), nefc(dd.nefc), jar(nefc), Ma(nv), grad(nv), search(nv), Mv(nv), jv(nefc), active(nefc) {}
  void mulM(const R* x, R* y) { for (int i = 0; i < nv; i++) { R s = 0; for (int k = 0; k < nv; k++) s += d.M[i * nv + k] * x[k]; y[i] = s; } }
  void mulJ(const R* x, R* y) { for (int i = 0; i < nefc; i++) { R s = 0; for (int k = 0; k < nv; k++) s += d.efc_J[i * nv + k] * x[k]; y[i] = s; } }
  // cost, force, active at jar (PrimalUpdateConstraint)
  R update_constraint(const std::vector<R>& ja, const R* qacc, const R* Ma_, bool write) {
    R c = 0;
    for (int i = 0; i < nefc; i++) {
      R x = ja[i], D = d.efc_D[i], Rr = d.efc_R[i], f = 0, a = 0, ci;
      if (d.efc_type[i] == EFC_FRICTION) {
        R fl = d.efc_floss[i];
        if (x <= -Rr * fl) { ci = fl * (-0.5 * Rr * fl - x); f = fl; }
        else if (x >= Rr * fl) { ci = fl * (-0.5 * Rr * fl + x); f = -fl; }
        else { ci = 0.5 * D * x * x; f = -D * x; a = 1; }
      } else {
        if (x < 0) { ci = 0.5 * D * x * x; f = -D * x; a = 1; } else ci = 0;
      }
      c += ci;
      if (write) { d.efc_force[i] = f; active[i] = a; }
    }
    R g = 0;
    for (int i = 0; i < nv; i++) g += (Ma_[i] - d.qfrc_smooth[i]) * (qacc[i] - d.qacc_smooth[i]);
    return c + 0.5 * g;
  }
  // derivative of the 1-D cost along `search` at alpha: returns (phi', phi'')
  void line_deriv(R alpha, R quad1, R quad2, R* d1, R* d2) {
    R g = quad1 + alpha * quad2, h = quad2;
    for (int i = 0; i < nefc; i++) {
      R x = jar[i] + alpha * jv[i], D = d.efc_D[i], Rr = d.efc_R[i];
      if (d.efc_type[i] == EFC_FRICTION) {
        R fl = d.efc_floss[i];
        if (x <= -Rr * fl) g += -fl * jv[i];
        else if (x >= Rr * fl) g += fl * jv[i];
        else { g += D * x * jv[i]; h += D * jv[i] * jv[i]; }
      } else if (x < 0) { g += D * x * jv[i]; h += D * jv[i] * jv[i]; }
    }
    *d1 = g; *d2 = h;
  }
};

void fwd_constraint(Data& d) {
  const Model& m = *d.m;
  int nv = m.nv, nefc = d.nefc;
  d.solver_niter = 0;
  if (nefc == 0) {
    d.qacc = d.qacc_smooth;
    d.qacc_warmstart = d.qacc_smooth;
    std::fill(d.qfrc_constraint.begin(), d.qfrc_constraint.end(), 0.0);
    return;
  }
  SolverCtx s(d);
  // warmstart selection
  std::vector<R> qacc(d.qacc_warmstart), tmpMa(nv), tmpjar(nefc);
  s.mulM(qacc.data(), tmpMa.data());
  s.mulJ(qacc.data(), tmpjar.data());
  for (int i = 0; i < nefc; i++) tmpjar[i] -= d.efc_aref[i];
  R cost_w = s.update_constraint(tmpjar, qacc.data(), tmpMa.data(), false);
  std::vector<R> Ma_s(d.qfrc_smooth), jar_s(nefc);
  s.mulJ(d.qacc_smooth.data(), jar_s.data());
  for (int i = 0; i < nefc; i++) jar_s[i] -= d.efc_aref[i];
  R cost_s = s.update_constraint(jar_s, d.qacc_smooth.data(), Ma_s.data(), false);
  if (!(cost_w < cost_s)) qacc = d.qacc_smooth;

  R scale = 1.0 / (m.meaninertia * std::max(1, nv));
  s.mulM(qacc.data(), s.Ma.data());
  s.mulJ(qacc.data(), s.jar.data());
  for (int i = 0; i < nefc; i++) s.jar[i] -= d.efc_aref[i];
  s.cost = s.update_constraint(s.jar, qacc.data(), s.Ma.data(), true);
  std::vector<R> H(nv * nv);
  auto gradient_and_search = [&]() {
    for (int i = 0; i < nv; i++) {
      R g = s.Ma[i] - d.qfrc_smooth[i];
      for (int r = 0; r < nefc; r++) g -= d.efc_J[r * nv + i] * d.efc_force[r];
      s.grad[i] = g;
    }
    H = d.M;
    for (int r = 0; r < nefc; r++) if (s.active[r] != 0) {
      const R* J = &d.efc_J[r * nv];
      for (int i = 0; i < nv; i++) if (J[i] != 0) for (int k = 0; k < nv; k++) H[i * nv + k] += d.efc_D[r] * J[i] * J[k];
    }
    cholesky(H, nv);
    for (int i = 0; i < nv; i++) s.search[i] = -s.grad[i];
    chol_solve(H, nv, s.search.data());
  };
  gradient_and_search();
  int iter = 0;
  while (iter < m.iterations) {
    // exact line search on the convex piecewise-quadratic 1-D cost (PrimalSearch)
    s.mulM(s.search.data(), s.Mv.data());
    s.mulJ(s.search.data(), s.jv.data());
    R quad1 = 0, quad2 = 0;
    for (int i = 0; i < nv; i++) { quad1 += s.search[i] * (s.Ma[i] - d.qfrc_smooth[i]); quad2 += s.search[i] * s.Mv[i]; }
    R lo = 0, hi = -1, alpha = 0, d1, d2;
    s.line_deriv(0, quad1, quad2, &d1, &d2);
    if (!(d1 < 0) || d2 <= 0) break;
    R d1_0 = d1;
    for (int ls = 0; ls < 100; ls++) {
      R next = alpha - d1 / d2;
      if (d1 < 0) lo = alpha; else hi = alpha;
      if (hi >= 0 && (next <= lo || next >= hi)) next = 0.5 * (lo + hi);
      alpha = next;
      s.line_deriv(alpha, quad1, quad2, &d1, &d2);
      if (std::fabs(d1) <= 1e-14 * std::fabs(d1_0)) break;
    }
    if (alpha == 0) break;
    for (int i = 0; i < nv; i++) { qacc[i] += alpha * s.search[i]; s.Ma[i] += alpha * s.Mv[i]; }
    for (int i = 0; i < nefc; i++) s.jar[i] += alpha * s.jv[i];
    R oldcost = s.cost;
    s.cost = s.update_constraint(s.jar, qacc.data(), s.Ma.data(), true);
    gradient_and_search();
    R improvement = scale * (oldcost - s.cost), gn = 0;
    for (int i = 0; i < nv; i++) gn += s.grad[i] * s.grad[i];
    R gradient = scale * std::sqrt(gn);
    iter++;
    if (improvement < m.tolerance || gradient < m.tolerance) break;
  }
  d.solver_niter = iter;
  d.qacc = qacc;
  d.qacc_warmstart = qacc;
  for (int i = 0; i < nv; i++) {
    R f = 0;
    for (int r = 0; r < nefc; r++) f += d.efc_J[r * nv + i] * d.efc_force[r];
    d.qfrc_constraint[i] = f;
  }
}

// ----------------------------------------------------------------------------- pipeline
void fwd_position(Data& d) {
  kinematics(d); com_pos(d); tendon(d); crb(d); collision(d);
}
void fwd_velocity(Data& d) { com_vel(d); passive(d); rne_bias(d); }
void step1(Data& d) { fwd_position(d); fwd_velocity(d); }
void quat_integrate(R* q, const R* w, R h) {  // mju_quatIntegrate
  R ax[3] = {w[0], w[1], w[2]};
  R n = norm3(ax);
  if (n < MINVAL) return;
  for (int k = 0; k < 3; k++) ax[k] /= n;
  R dq[4], r[4];
  axisangle2quat(dq, ax, h * n);
  quat_mul(r, q, dq);
  quat_normalize(r);
  std::memcpy(q, r, sizeof(r));
}
void euler(Data& d) {  // mj_Euler with implicit joint damping
  const Model& m = *d.m;
  int nv = m.nv;
  const R* damp = m.F(DEXF_DOF_DAMPING);
  std::vector<R> qacc(d.qacc);
  bool any = false;
  for (int i = 0; i < nv; i++) any |= damp[i] > 0;
  if (any) {
    std::vector<R> A(d.M);
    for (int i = 0; i < nv; i++) A[i * nv + i] += m.timestep * damp[i];
    cholesky(A, nv);
    for (int i = 0; i < nv; i++) qacc[i] = d.qfrc_smooth[i] + d.qfrc_constraint[i];
    chol_solve(A, nv, qacc.data());
  }
  for (int i = 0; i < nv; i++) d.qvel[i] += m.timestep * qacc[i];
  const int32_t *jtype = m.I(DEXF_JNT_TYPE), *jq = m.I(DEXF_JNT_QPOSADR), *jd = m.I(DEXF_JNT_DOFADR);
  for (int j = 0; j < m.njnt; j++) {
    int qa = jq[j], da = jd[j];
    if (jtype[j] == 0) {
      for (int k = 0; k < 3; k++) d.qpos[qa + k] += m.timestep * d.qvel[da + k];
      quat_integrate(&d.qpos[qa + 3], &d.qvel[da + 3], m.timestep);
    } else {
      d.qpos[qa] += m.timestep * d.qvel[da];
    }
  }
  d.time += m.timestep;
}
void step2(Data& d) {
  fwd_actuation(d); fwd_acceleration(d); make_constraint(d); fwd_constraint(d); euler(d);
}
void forward(Data& d) {
  step1(d); fwd_actuation(d); fwd_acceleration(d); make_constraint(d); fwd_constraint(d);
}
void step(Data& d) { step1(d); step2(d); }

}  // namespace

// ============================================================================= C API (ctypes; test infrastructure)
extern "C" {

void* orc_model_create(const void* blob, uint64_t nbytes) {
  const DexBlobHeader* h = (const DexBlobHeader*)blob;
  if (nbytes < sizeof(DexBlobHeader) || h->magic != DEX_BLOB_MAGIC || h->version != DEX_BLOB_VERSION ||
      h->nfields != DEXF_COUNT || h->total_bytes != nbytes) return nullptr;
  Model* m = new Model;
  m->blob.assign((const uint8_t*)blob, (const uint8_t*)blob + nbytes);
  const int32_t* D = m->I(DEXF_DIMS);
  m->nbody = D[DEXD_NBODY]; m->njnt = D[DEXD_NJNT]; m->nq = D[DEXD_NQ]; m->nv = D[DEXD_NV]; m->nu = D[DEXD_NU];
  m->ntendon = D[DEXD_NTENDON]; m->ngeom = D[DEXD_NGEOM]; m->npair = D[DEXD_NPAIR]; m->nsite = D[DEXD_NSITE];
  m->nconmax = D[DEXD_NCONMAX]; m->iterations = D[DEXD_SOLVER_ITERATIONS]; m->mpr_iterations = D[DEXD_MPR_ITERATIONS];
  const R* O = m->F(DEXF_OPT);
  m->timestep = O[DEXO_TIMESTEP]; m->gravity[0] = O[DEXO_GRAVITY_X]; m->gravity[1] = O[DEXO_GRAVITY_Y];
  m->gravity[2] = O[DEXO_GRAVITY_Z]; m->tolerance = O[DEXO_TOLERANCE]; m->mpr_tolerance = O[DEXO_MPR_TOLERANCE];
  m->meaninertia = O[DEXO_MEANINERTIA];
  return m;
}
void orc_model_destroy(void* m) { delete (Model*)m; }
void* orc_data_create(void* m) { return new Data((Model*)m); }
void orc_data_destroy(void* d) { delete (Data*)d; }

void orc_set_state(void* dp, const double* qpos, const double* qvel, const double* ctrl, const double* warm, double time) {
  Data& d = *(Data*)dp;
  if (qpos) d.qpos.assign(qpos, qpos + d.m->nq);
  if (qvel) d.qvel.assign(qvel, qvel + d.m->nv);
  if (ctrl) d.ctrl.assign(ctrl, ctrl + d.m->nu);
  if (warm) d.qacc_warmstart.assign(warm, warm + d.m->nv);
  d.time = time;
}
void orc_get_state(void* dp, double* qpos, double* qvel, double* warm, double* time) {
  Data& d = *(Data*)dp;
  if (qpos) std::copy(d.qpos.begin(), d.qpos.end(), qpos);
  if (qvel) std::copy(d.qvel.begin(), d.qvel.end(), qvel);
  if (warm) std::copy(d.qacc_warmstart.begin(), d.qacc_warmstart.end(), warm);
  if (time) *time = d.time;
}
void orc_forward(void* d) { forward(*(Data*)d); }
void orc_step(void* d) { step(*(Data*)d); }
void orc_step1(void* d) { step1(*(Data*)d); }
void orc_step2(void* d) { step2(*(Data*)d); }
void orc_kinematics(void* dp) { Data& d = *(Data*)dp; kinematics(d); com_pos(d); }

// field getter by name (doubles); returns element count, -1 if unknown
int64_t orc_get(void* dp, const char* name, double* out, int64_t cap) {
  Data& d = *(Data*)dp;
  const std::vector<R>* v = nullptr;
  std::string n(name);
#define F(x) if (n == #x) v = &d.x;
  F(xpos) F(xquat) F(xmat) F(xipos) F(ximat) F(xanchor) F(xaxis) F(geom_xpos) F(geom_xmat) F(site_xpos) F(site_xmat)
  F(subtree_com) F(cinert) F(cdof) F(cvel) F(cdof_dot) F(M) F(qfrc_bias) F(qfrc_passive) F(qfrc_actuator)
  F(qfrc_applied) F(qfrc_smooth) F(qacc_smooth) F(ten_length) F(act_force) F(efc_J) F(efc_pos) F(efc_margin)
  F(efc_D) F(efc_R) F(efc_aref) F(efc_force) F(efc_floss) F(qacc) F(qfrc_constraint) F(qpos) F(qvel) F(ctrl) F(qacc_warmstart)
#undef F
  if (!v) return -1;
  int64_t c = (int64_t)v->size();
  if (out) std::copy(v->begin(), v->begin() + std::min(c, cap), out);
  return c;
}
int orc_nefc(void* d) { return ((Data*)d)->nefc; }
int orc_solver_niter(void* d) { return ((Data*)d)->solver_niter; }
int orc_efc_types(void* dp, int* type, int* id, int cap) {
  Data& d = *(Data*)dp;
  for (int i = 0; i < d.nefc && i < cap; i++) { type[i] = d.efc_type[i]; id[i] = d.efc_id[i]; }
  return d.nefc;
}
// contacts: geoms[2*i], vals[i*13] = dist, pos(3), frame(9)
int orc_contacts(void* dp, int* geoms, double* vals, int cap) {
  Data& d = *(Data*)dp;
  int n = (int)d.contacts.size();
  for (int i = 0; i < n && i < cap; i++) {
    const Contact& c = d.contacts[i];
    geoms[2 * i] = c.geom1; geoms[2 * i + 1] = c.geom2;
    vals[13 * i] = c.dist;
    std::memcpy(vals + 13 * i + 1, c.pos, 3 * sizeof(R));
    std::memcpy(vals + 13 * i + 4, c.frame, 9 * sizeof(R));
  }
  return n;
}
// mj_jac at a world point attached to body → jacp, jacr (3 x nv each)
void orc_jac(void* dp, int body, const double* point, double* jacp, double* jacr) { jac(*(Data*)dp, jacp, jacr, point, body); }
// mj_objectVelocity(flg_local = 0) at a world point attached to body: out = [ang(3), lin(3)]
void orc_object_velocity(void* dp, int body, const double* point, double* out) {
  Data& d = *(Data*)dp;
  const int32_t* root = d.m->I(DEXF_BODY_ROOT);
  const R* cv = &d.cvel[6 * body];
  R dif[3], t[3];
  for (int k = 0; k < 3; k++) dif[k] = point[k] - d.subtree_com[3 * root[body] + k];
  cross(t, cv, dif);
  for (int k = 0; k < 3; k++) { out[k] = cv[k]; out[3 + k] = cv[3 + k] + t[k]; }
}
// MPR between two geoms at the current kinematics (for narrowphase unit tests)
int orc_mpr(void* dp, int g1, int g2, double* depth, double* dir, double* pos) {
  Data& d = *(Data*)dp;
  Shape a{&d, g1, 0}, b{&d, g2, 0};
  return mpr_penetration(a, b, d.m->mpr_iterations, d.m->mpr_tolerance, depth, dir, pos);
}
void orc_support(void* dp, int g, const double* dir, double* out) { Shape s{(Data*)dp, g, 0}; support(s, dir, out); }

// Batched stepping over `n` independent envs on `nthreads` host threads (CPU baseline).
// State arrays are [n][nq] / [n][nv] / [n][nu]; each env runs `nsub` mj_steps followed by the
// position/velocity refresh dm_control's legacy step performs (mj_step2 ... mj_step1).
void orc_step_batch(void* mp, int64_t n, int nsub, double* qpos, double* qvel, const double* ctrl, double* warm,
                    int* ncon_out, int nthreads) {
  Model* m = (Model*)mp;
  auto work = [&](int64_t lo, int64_t hi) {
    Data d(m);
    for (int64_t e = lo; e < hi; e++) {
      orc_set_state(&d, qpos + e * m->nq, qvel + e * m->nv, ctrl + e * m->nu, warm + e * m->nv, 0);
      for (int s = 0; s < nsub; s++) step(d);
      step1(d);
      orc_get_state(&d, qpos + e * m->nq, qvel + e * m->nv, warm + e * m->nv, nullptr);
      if (ncon_out) ncon_out[e] = (int)d.contacts.size();
    }
  };
  if (nthreads <= 1) { work(0, n); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) th.emplace_back(work, n * t / nthreads, n * (t + 1) / nthreads);
  for (auto& t : th) t.join();
}

}  // extern "C"
