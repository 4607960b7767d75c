// dex_collide.cuh — K1: warp-per-env collision detection (restates mj_collision for the
// suite's geom pairs).
//
//   broadphase  : static candidate pairs grouped by body pair (compile-time filters already
//                 applied: same weld, parent-child, <exclude>, contype/conaffinity), culled at
//                 run time by body bounding spheres then geom bounding spheres, 32 tests per
//                 warp instruction, survivors compacted with ballots.
//   narrowphase : plane-box (mjc_PlaneBox), plane-mesh (mjc_PlaneConvex) and convex-convex
//                 through libccd's MPR (ccdMPRPenetration) with mpr_tolerance/mpr_iterations.
//                 The portal state is warp-uniform; the mesh support function is the parallel
//                 part (lanes scan hull vertices, warp argmax with lowest-index tie-break).
// Contacts are written compactly to global memory [env][DEX_MAXCON][8] for K2.
#ifndef DEX_COLLIDE_CUH_
#define DEX_COLLIDE_CUH_

#include "dex_kin.cuh"

// per-warp shared-memory footprint of K1, in units of T
#define DEX_K1_SMEM_T (DEX_MAXBODY * 16 + 48 + DEX_MAXCAND + 8 * 4)

template <typename T>
struct CollideCtx {
  const DevModel<T>* m;
  const T* xpos;
  const T* xmat;
  T half_margin;  // mjc_Convex inflates both shapes by margin/2 along the query direction
};

// world-frame support point of geom g in unit direction dir (mjccd_support, margin 0)
template <typename T>
DEX_DEV void dex_support(const CollideCtx<T>& c, int g, int gtype, const T* dir, T* out) {
  const DevModel<T>& m = *c.m;
  const int b = dex_ldg(m.geom_body + g);
  const T* R = c.xmat + 9 * b;
  T ld[3], loc[3];
  rotT_vec(ld, R, dir);
  if (gtype == DEX_GEOM_MESH) {
    const int adr = dex_ldg(m.geom_vertadr + g), n = dex_ldg(m.geom_vertnum + g);
    const T* v = m.vert + 3 * adr;
    T best = -Real<T>::big();
    int bi = 0x7fffffff;
    for (int i = dex_lane(); i < n; i += 32) {
      T t = dex_ldg(v + 3 * i) * ld[0] + dex_ldg(v + 3 * i + 1) * ld[1] + dex_ldg(v + 3 * i + 2) * ld[2];
      if (t > best) { best = t; bi = i; }
    }
    dex_warp_argmax(best, bi);
    loc[0] = dex_ldg(v + 3 * bi); loc[1] = dex_ldg(v + 3 * bi + 1); loc[2] = dex_ldg(v + 3 * bi + 2);
  } else {  // box
    const T* gm = m.geom_mat + 9 * g;
    T G[9];
    for (int i = 0; i < 9; ++i) G[i] = dex_ldg(gm + i);
    T l2[3], cn[3];
    rotT_vec(l2, G, ld);
    cn[0] = l2[0] > 0 ? dex_ldg(m.geom_size + 3 * g) : -dex_ldg(m.geom_size + 3 * g);
    cn[1] = l2[1] > 0 ? dex_ldg(m.geom_size + 3 * g + 1) : -dex_ldg(m.geom_size + 3 * g + 1);
    cn[2] = l2[2] > 0 ? dex_ldg(m.geom_size + 3 * g + 2) : -dex_ldg(m.geom_size + 3 * g + 2);
    rot_vec(loc, G, cn);
    loc[0] += dex_ldg(m.geom_pos + 3 * g); loc[1] += dex_ldg(m.geom_pos + 3 * g + 1); loc[2] += dex_ldg(m.geom_pos + 3 * g + 2);
  }
  rot_vec(out, R, loc);
  out[0] += c.xpos[3 * b]; out[1] += c.xpos[3 * b + 1]; out[2] += c.xpos[3 * b + 2];
}

template <typename T>
DEX_DEV void dex_geom_center(const CollideCtx<T>& c, int g, T* out) {
  const DevModel<T>& m = *c.m;
  const int b = dex_ldg(m.geom_body + g);
  T gp[3] = {dex_ldg(m.geom_pos + 3 * g), dex_ldg(m.geom_pos + 3 * g + 1), dex_ldg(m.geom_pos + 3 * g + 2)};
  rot_vec(out, c.xmat + 9 * b, gp);
  out[0] += c.xpos[3 * b]; out[1] += c.xpos[3 * b + 1]; out[2] += c.xpos[3 * b + 2];
}

// true if the sphere (world centre c, radius r) may touch the body-frame AABB of geom g
template <typename T>
DEX_DEV bool dex_sphere_vs_aabb(const CollideCtx<T>& cx, const T* c, T r, int g) {
  const DevModel<T>& m = *cx.m;
  const int b = dex_ldg(m.geom_body + g);
  T d[3] = {c[0] - cx.xpos[3 * b], c[1] - cx.xpos[3 * b + 1], c[2] - cx.xpos[3 * b + 2]}, l[3];
  rotT_vec(l, cx.xmat + 9 * b, d);
  T s = 0;
  for (int k = 0; k < 3; ++k) {
    T e = dx_abs(l[k] - dex_ldg(m.geom_aabb + 6 * g + k)) - dex_ldg(m.geom_aabb + 6 * g + 3 + k);
    if (e > 0) s += e * e;
  }
  return s <= r * r;
}

// ---- libccd helpers ---------------------------------------------------------------------
// libccd's "is zero" tests are ABSOLUTE thresholds at CCD_EPS = DBL_EPSILON (MuJoCo builds libccd
// in double).  They guard exact degeneracies, so the FP32 build keeps the same absolute value
// (an FP32 epsilon would wrongly fire on centimetre-scale cross products); only the relative
// comparison in ccdEq uses the working precision's epsilon.
template <typename T> DEX_DEV bool ccd_zero(T x) { return dx_abs(x) < T(2.220446049250313e-16); }
template <typename T> DEX_DEV bool ccd_eq(T a, T b) {
  T ab = dx_abs(a - b);
  if (ab < T(2.220446049250313e-16)) return true;
  T fa = dx_abs(a), fb = dx_abs(b);
  return fb > fa ? ab < Real<T>::eps() * fb : ab < Real<T>::eps() * fa;
}
template <typename T> struct SupPt { T v[3], v1[3], v2[3]; };

template <typename T>
DEX_DEV void dex_mink(const CollideCtx<T>& c, int g1, int t1, int g2, int t2, const T* dir, SupPt<T>& p) {
  T nd[3] = {-dir[0], -dir[1], -dir[2]};
  dex_support(c, g1, t1, dir, p.v1);
  dex_support(c, g2, t2, nd, p.v2);
  for (int k = 0; k < 3; ++k) { p.v1[k] += dir[k] * c.half_margin; p.v2[k] -= dir[k] * c.half_margin; }
  p.v[0] = p.v1[0] - p.v2[0]; p.v[1] = p.v1[1] - p.v2[1]; p.v[2] = p.v1[2] - p.v2[2];
}
template <typename T> DEX_DEV T point_seg_dist2(const T* x0, const T* b, T* wit) {  // P = origin
  T dd[3] = {b[0] - x0[0], b[1] - x0[1], b[2] - x0[2]};
  T t = -dot3(x0, dd) / dot3(dd, dd);
  if (t < 0 || ccd_zero(t)) { wit[0] = x0[0]; wit[1] = x0[1]; wit[2] = x0[2]; }
  else if (t > 1 || ccd_eq(t, T(1))) { wit[0] = b[0]; wit[1] = b[1]; wit[2] = b[2]; }
  else { wit[0] = x0[0] + t * dd[0]; wit[1] = x0[1] + t * dd[1]; wit[2] = x0[2] + t * dd[2]; }
  return dot3(wit, wit);
}
template <typename T> DEX_DEV T point_tri_dist2(const T* x0, const T* B, const T* C, T* wit) {  // P = origin
  T d1[3] = {B[0] - x0[0], B[1] - x0[1], B[2] - x0[2]};
  T d2[3] = {C[0] - x0[0], C[1] - x0[1], C[2] - x0[2]};
  T v = dot3(d1, d1), w = dot3(d2, d2), p = dot3(x0, d1), q = dot3(x0, d2), r = dot3(d1, d2);
  T s = (q * r - w * p) / (w * v - r * r);
  T t = (-s * r - q) / w;
  if ((ccd_zero(s) || s > 0) && (ccd_eq(s, T(1)) || s < 1) && (ccd_zero(t) || t > 0) && (ccd_eq(t, T(1)) || t < 1) &&
      (ccd_eq(t + s, T(1)) || t + s < 1)) {
    wit[0] = x0[0] + s * d1[0] + t * d2[0];
    wit[1] = x0[1] + s * d1[1] + t * d2[1];
    wit[2] = x0[2] + s * d1[2] + t * d2[2];
    return dot3(wit, wit);
  }
  T w2[3];
  T dist = point_seg_dist2(x0, B, wit);
  T d = point_seg_dist2(x0, C, w2);
  if (d < dist) { dist = d; wit[0] = w2[0]; wit[1] = w2[1]; wit[2] = w2[2]; }
  d = point_seg_dist2(B, C, w2);
  if (d < dist) { dist = d; wit[0] = w2[0]; wit[1] = w2[1]; wit[2] = w2[2]; }
  return dist;
}
template <typename T> DEX_DEV void portal_dir(const SupPt<T>* s, T* dir) {
  T a[3] = {s[2].v[0] - s[1].v[0], s[2].v[1] - s[1].v[1], s[2].v[2] - s[1].v[2]};
  T b[3] = {s[3].v[0] - s[1].v[0], s[3].v[1] - s[1].v[1], s[3].v[2] - s[1].v[2]};
  cross3(dir, a, b);
  normalize3(dir);
}
template <typename T> DEX_DEV bool portal_reach_tol(const SupPt<T>* s, const SupPt<T>& v4, const T* dir, T tol) {
  T dv4 = dot3(v4.v, dir);
  T mn = dx_min(dv4 - dot3(s[1].v, dir), dx_min(dv4 - dot3(s[2].v, dir), dv4 - dot3(s[3].v, dir)));
  return ccd_eq(mn, tol) || mn < tol;
}
template <typename T> DEX_DEV void expand_portal(SupPt<T>* s, const SupPt<T>& v4) {
  T x[3];
  cross3(x, v4.v, s[0].v);
  if (dot3(s[1].v, x) > 0) {
    if (dot3(s[2].v, x) > 0) s[1] = v4; else s[3] = v4;
  } else {
    if (dot3(s[3].v, x) > 0) s[2] = v4; else s[1] = v4;
  }
}

// ccdMPRPenetration.  All lanes run this with identical (warp-uniform) state.
// returns true on intersection; depth/dir/pos as libccd.
template <typename T>
DEX_DEV bool dex_mpr(const CollideCtx<T>& c, int g1, int t1, int g2, int t2, T* depth, T* pdir, T* pos) {
  const DevModel<T>& m = *c.m;
  SupPt<T> s[4], v4;
  T dir[3], va[3], vb[3];
  dex_geom_center(c, g1, s[0].v1);
  dex_geom_center(c, g2, s[0].v2);
  for (int k = 0; k < 3; ++k) s[0].v[k] = s[0].v1[k] - s[0].v2[k];
  if (ccd_zero(s[0].v[0]) && ccd_zero(s[0].v[1]) && ccd_zero(s[0].v[2])) s[0].v[0] += T(2.220446049250313e-15);
  dir[0] = -s[0].v[0]; dir[1] = -s[0].v[1]; dir[2] = -s[0].v[2];
  normalize3(dir);
  dex_mink(c, g1, t1, g2, t2, dir, s[1]);
  T dt = dot3(s[1].v, dir);
  if (ccd_zero(dt) || dt < 0) return false;
  cross3(dir, s[0].v, s[1].v);
  if (ccd_zero(dot3(dir, dir))) {
    if (ccd_zero(s[1].v[0]) && ccd_zero(s[1].v[1]) && ccd_zero(s[1].v[2])) {  // touching: findPenetrTouch
      *depth = 0; pdir[0] = pdir[1] = pdir[2] = 0;
    } else {  // origin on v0-v1: findPenetrSegment
      *depth = dx_sqrt(dot3(s[1].v, s[1].v));
      pdir[0] = s[1].v[0]; pdir[1] = s[1].v[1]; pdir[2] = s[1].v[2];
      normalize3(pdir);
    }
    for (int k = 0; k < 3; ++k) pos[k] = T(0.5) * (s[1].v1[k] + s[1].v2[k]);
    return true;
  }
  normalize3(dir);
  dex_mink(c, g1, t1, g2, t2, dir, s[2]);
  dt = dot3(s[2].v, dir);
  if (ccd_zero(dt) || dt < 0) return false;
  for (int k = 0; k < 3; ++k) { va[k] = s[1].v[k] - s[0].v[k]; vb[k] = s[2].v[k] - s[0].v[k]; }
  cross3(dir, va, vb);
  normalize3(dir);
  if (dot3(dir, s[0].v) > 0) {
    SupPt<T> tmp = s[1]; s[1] = s[2]; s[2] = tmp;
    dir[0] = -dir[0]; dir[1] = -dir[1]; dir[2] = -dir[2];
  }
  for (int guard = 0;; ++guard) {  // discoverPortal: find v3
    if (guard > 100) return false;
    dex_mink(c, g1, t1, g2, t2, dir, s[3]);
    dt = dot3(s[3].v, dir);
    if (ccd_zero(dt) || dt < 0) return false;
    bool cont = false;
    cross3(va, s[1].v, s[3].v);
    dt = dot3(va, s[0].v);
    if (dt < 0 && !ccd_zero(dt)) { s[2] = s[3]; cont = true; }
    if (!cont) {
      cross3(va, s[3].v, s[2].v);
      dt = dot3(va, s[0].v);
      if (dt < 0 && !ccd_zero(dt)) { s[1] = s[3]; cont = true; }
    }
    if (!cont) break;
    for (int k = 0; k < 3; ++k) { va[k] = s[1].v[k] - s[0].v[k]; vb[k] = s[2].v[k] - s[0].v[k]; }
    cross3(dir, va, vb);
    normalize3(dir);
  }
  for (int guard = 0;; ++guard) {  // refinePortal
    if (guard > 100) return false;
    portal_dir(s, dir);
    dt = dot3(dir, s[1].v);
    if (ccd_zero(dt) || dt > 0) break;
    dex_mink(c, g1, t1, g2, t2, dir, v4);
    dt = dot3(v4.v, dir);
    if (!(ccd_zero(dt) || dt > 0) || portal_reach_tol(s, v4, dir, m.mpr_tolerance)) return false;
    expand_portal(s, v4);
  }
  for (int it = 0;; ++it) {  // findPenetr
    portal_dir(s, dir);
    dex_mink(c, g1, t1, g2, t2, dir, v4);
    if (portal_reach_tol(s, v4, dir, m.mpr_tolerance) || it > m.mpr_iterations) {
      T d2 = point_tri_dist2(s[1].v, s[2].v, s[3].v, pdir);
      *depth = dx_sqrt(d2);
      if (ccd_zero(pdir[0]) && ccd_zero(pdir[1]) && ccd_zero(pdir[2])) { pdir[0] = dir[0]; pdir[1] = dir[1]; pdir[2] = dir[2]; }
      normalize3(pdir);
      // findPos: barycentric coordinates of the origin in the portal tetrahedron
      T b[4], t[3];
      cross3(t, s[1].v, s[2].v); b[0] = dot3(t, s[3].v);
      cross3(t, s[3].v, s[2].v); b[1] = dot3(t, s[0].v);
      cross3(t, s[0].v, s[1].v); b[2] = dot3(t, s[3].v);
      cross3(t, s[2].v, s[1].v); b[3] = dot3(t, s[0].v);
      T sum = b[0] + b[1] + b[2] + b[3];
      if (ccd_zero(sum) || sum < 0) {
        b[0] = 0;
        cross3(t, s[2].v, s[3].v); b[1] = dot3(t, dir);
        cross3(t, s[3].v, s[1].v); b[2] = dot3(t, dir);
        cross3(t, s[1].v, s[2].v); b[3] = dot3(t, dir);
        sum = b[1] + b[2] + b[3];
      }
      T inv = T(1) / sum;
      for (int k = 0; k < 3; ++k) {
        T p1 = b[0] * s[0].v1[k] + b[1] * s[1].v1[k] + b[2] * s[2].v1[k] + b[3] * s[3].v1[k];
        T p2 = b[0] * s[0].v2[k] + b[1] * s[1].v2[k] + b[2] * s[2].v2[k] + b[3] * s[3].v2[k];
        pos[k] = T(0.5) * (p1 + p2) * inv;
      }
      return true;
    }
    expand_portal(s, v4);
  }
}

// ------------------------------------------------------------------------------------------
// K1 body.  smem: [xpos | xquat | xmat | qpos | cand(int)]
// ------------------------------------------------------------------------------------------
template <typename T>
DEX_DEV void dex_store_contact(T* dst, T dist, const T* pos, const T* normal, int pair) {
  dst[0] = dist; dst[1] = pos[0]; dst[2] = pos[1]; dst[3] = pos[2];
  dst[4] = normal[0]; dst[5] = normal[1]; dst[6] = normal[2];
  dst[7] = (T)pair;  // exact for pair ids < 2^24
}

template <typename T>
DEX_DEV void dex_collide_env(const DevModel<T>& m, const BatchState<T>& st, int env, T* smem) {
  if (st.skip && st.skip[env]) return;  // warp-uniform
  const int lane = dex_lane();
  T* xpos = smem;
  T* xquat = xpos + DEX_MAXBODY * 3;
  T* xmat = xquat + DEX_MAXBODY * 4;
  T* qpos = xmat + DEX_MAXBODY * 9;           // 48
  int* cand = (int*)(qpos + 48);              // DEX_MAXCAND ints (fits: sizeof(int) <= sizeof(T))
  for (int i = lane; i < m.nq; i += 32) qpos[i] = st.qpos[(size_t)env * st.qpos_stride + i];
  dex_syncwarp();
  KinBuf<T> kb = {xpos, xquat, xmat, nullptr, nullptr};
  dex_fk(m, qpos, kb);
  CollideCtx<T> cx = {&m, xpos, xmat, T(0)};

  // ---- broadphase
  int ncand = 0;
  for (int bp0 = 0; bp0 < m.nbodypair; bp0 += 32) {
    const int bp = bp0 + lane;
    bool pass = false;
    if (bp < m.nbodypair) {
      const int b1 = dex_ldg(m.bp_body1 + bp), b2 = dex_ldg(m.bp_body2 + bp);
      const T r1 = dex_ldg(m.body_bsphere + 4 * b1 + 3), r2 = dex_ldg(m.body_bsphere + 4 * b2 + 3);
      if (r1 < 0 || r2 < 0) pass = true;  // body holds a plane: resolved per geom
      else {
        T c1[3], c2[3];
        T l1[3] = {dex_ldg(m.body_bsphere + 4 * b1), dex_ldg(m.body_bsphere + 4 * b1 + 1), dex_ldg(m.body_bsphere + 4 * b1 + 2)};
        T l2[3] = {dex_ldg(m.body_bsphere + 4 * b2), dex_ldg(m.body_bsphere + 4 * b2 + 1), dex_ldg(m.body_bsphere + 4 * b2 + 2)};
        rot_vec(c1, xmat + 9 * b1, l1);
        rot_vec(c2, xmat + 9 * b2, l2);
        T d[3] = {c2[0] + xpos[3 * b2] - c1[0] - xpos[3 * b1], c2[1] + xpos[3 * b2 + 1] - c1[1] - xpos[3 * b1 + 1],
                  c2[2] + xpos[3 * b2 + 2] - c1[2] - xpos[3 * b1 + 2]};
        T bound = r1 + r2;
        pass = dot3(d, d) <= bound * bound;
      }
    }
    unsigned mask = dex_ballot(pass);
    while (mask) {
      const int kk = dex_ffs(mask) - 1;
      mask &= mask - 1;
      const int bpk = bp0 + kk;
      const int adr = dex_ldg(m.bp_pairadr + bpk), num = dex_ldg(m.bp_pairnum + bpk);
      for (int p0 = 0; p0 < num; p0 += 32) {
        const int p = adr + p0 + lane;
        bool ok = false;
        if (p0 + lane < num) {
          const int g1 = dex_ldg(m.pair_geom1 + p), g2 = dex_ldg(m.pair_geom2 + p);
          const T margin = dex_ldg(m.pair_margin + p);
          T c1[3], c2[3];
          dex_geom_center(cx, g1, c1);
          dex_geom_center(cx, g2, c2);
          T d[3] = {c2[0] - c1[0], c2[1] - c1[1], c2[2] - c1[2]};
          if (dex_ldg(m.geom_type + g1) == DEX_GEOM_PLANE) {
            const int b1 = dex_ldg(m.geom_body + g1);
            const T* R = xmat + 9 * b1;
            const T* G = m.geom_mat + 9 * g1;
            T nl[3] = {dex_ldg(G + 2), dex_ldg(G + 5), dex_ldg(G + 8)}, n[3];
            rot_vec(n, R, nl);
            ok = dot3(d, n) <= margin + dex_ldg(m.geom_rbound + g2);
          } else {
            const T r1 = dex_ldg(m.geom_rbound + g1), r2 = dex_ldg(m.geom_rbound + g2);
            T bound = margin + r1 + r2;
            ok = dot3(d, d) <= bound * bound;
            // conservative refinements: bounding sphere of one geom against the body-frame AABB of the other
            if (ok) ok = dex_sphere_vs_aabb(cx, c2, r2 + margin, g1);
            if (ok) ok = dex_sphere_vs_aabb(cx, c1, r1 + margin, g2);
          }
        }
        const unsigned m2 = dex_ballot(ok);
        if (ok) {
          const int idx = ncand + dex_popc(m2 & ((1u << lane) - 1u));
          if (idx < DEX_MAXCAND) cand[idx] = p;
        }
        ncand += dex_popc(m2);
      }
    }
  }
  unsigned status = 0;
  if (ncand > DEX_MAXCAND) { status |= DEX_STATUS_CAND_OVERFLOW; ncand = DEX_MAXCAND; }
  dex_syncwarp();

  // ---- narrowphase
  T* cout = st.contact + (size_t)env * DEX_MAXCON * DEX_CON_STRIDE;
  int ncon = 0;
  for (int ci = 0; ci < ncand; ++ci) {
    const int p = cand[ci];
    const int g1 = dex_ldg(m.pair_geom1 + p), g2 = dex_ldg(m.pair_geom2 + p);
    const int t1 = dex_ldg(m.geom_type + g1), t2 = dex_ldg(m.geom_type + g2);
    const T margin = dex_ldg(m.pair_margin + p);
    if (t1 == DEX_GEOM_PLANE) {
      const int b1 = dex_ldg(m.geom_body + g1), b2 = dex_ldg(m.geom_body + g2);
      const T* G1 = m.geom_mat + 9 * g1;
      T nl[3] = {dex_ldg(G1 + 2), dex_ldg(G1 + 5), dex_ldg(G1 + 8)}, n[3], p1[3];
      rot_vec(n, xmat + 9 * b1, nl);
      dex_geom_center(cx, g1, p1);
      if (t2 == DEX_GEOM_BOX) {  // mjc_PlaneBox: lanes 0..7 test the corners, first 4 hits kept
        T c2[3];
        dex_geom_center(cx, g2, c2);
        T dif[3] = {c2[0] - p1[0], c2[1] - p1[1], c2[2] - p1[2]};
        const T dist = dot3(dif, n);
        T corner[3] = {0, 0, 0};
        T ldist = 0;
        bool hit = false;
        if (lane < 8) {
          T sz[3] = {dex_ldg(m.geom_size + 3 * g2), dex_ldg(m.geom_size + 3 * g2 + 1), dex_ldg(m.geom_size + 3 * g2 + 2)};
          T vec[3] = {(lane & 1) ? sz[0] : -sz[0], (lane & 2) ? sz[1] : -sz[1], (lane & 4) ? sz[2] : -sz[2]};
          T G[9], tmp[3];
          for (int i = 0; i < 9; ++i) G[i] = dex_ldg(m.geom_mat + 9 * g2 + i);
          rot_vec(tmp, G, vec);
          rot_vec(corner, xmat + 9 * b2, tmp);
          ldist = dot3(n, corner);
          hit = !(dist + ldist > margin || ldist > 0);
        }
        const unsigned hm = dex_ballot(hit);
        const int rank = dex_popc(hm & ((1u << lane) - 1u));
        if (hit && rank < 4) {
          if (ncon + rank < DEX_MAXCON) {
            T cd = dist + ldist;
            T cp[3] = {corner[0] + c2[0] - n[0] * cd * T(0.5), corner[1] + c2[1] - n[1] * cd * T(0.5),
                       corner[2] + c2[2] - n[2] * cd * T(0.5)};
            dex_store_contact(cout + (ncon + rank) * DEX_CON_STRIDE, cd, cp, n, p);
          }
        }
        int add = dex_popc(hm); if (add > 4) add = 4;
        if (ncon + add > DEX_MAXCON) { status |= DEX_STATUS_CON_OVERFLOW; add = DEX_MAXCON - ncon; }
        ncon += add;
      } else if (t2 == DEX_GEOM_MESH) {  // mjc_PlaneConvex: deepest vertex (+ up to 3 hull neighbours: TODO round 2)
        T nn[3] = {-n[0], -n[1], -n[2]}, w[3];
        dex_support(cx, g2, t2, nn, w);
        T dif[3] = {w[0] - p1[0], w[1] - p1[1], w[2] - p1[2]};
        const T dist = dot3(dif, n);
        if (dist <= margin) {
          if (ncon < DEX_MAXCON) {
            T cp[3] = {w[0] - T(0.5) * dist * n[0], w[1] - T(0.5) * dist * n[1], w[2] - T(0.5) * dist * n[2]};
            if (lane == 0) dex_store_contact(cout + ncon * DEX_CON_STRIDE, dist, cp, n, p);
            ncon++;
          } else status |= DEX_STATUS_CON_OVERFLOW;
        }
      }
    } else {
      T depth, dir[3], pos[3];
      cx.half_margin = T(0.5) * margin;
      if (dex_mpr(cx, g1, t1, g2, t2, &depth, dir, pos) && !(dir[0] == 0 && dir[1] == 0 && dir[2] == 0)) {
        if (ncon < DEX_MAXCON) {
          if (lane == 0) dex_store_contact(cout + ncon * DEX_CON_STRIDE, margin - depth, pos, dir, p);
          ncon++;
        } else status |= DEX_STATUS_CON_OVERFLOW;
      }
    }
  }
  if (lane == 0) {
    st.ncon[env] = ncon;
    if (status) st.status[env] |= status;
  }
}

#endif  // DEX_COLLIDE_CUH_
