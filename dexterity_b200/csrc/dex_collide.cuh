// dex_collide.cuh — collision detection (restates mj_collision for the suite's geom pairs,
// SURVEY.md §8a row 14) as three kernels:
//
//   K1a broadphase, WARP per env: static candidate pairs grouped by body pair (compile-time filters
//       already applied: same weld, parent-child, <exclude>, contype/conaffinity).  Run-time culls are
//       conservative and therefore do not change the contact set: body bounding spheres (plane:
//       half-space test) and body OBBs, then per geom pair, in the frame of body A so that no per-geom
//       world transform is needed: bounding spheres, sphere vs body-frame AABB, AABB vs AABB as
//       oriented boxes.  32 tests per warp instruction; survivors are compacted with ballots.  Plane
//       pairs (mjc_PlaneBox, mjc_PlaneConvex) are finished here; every convex-convex survivor gets a
//       slot in the env's staging list and one entry in a GLOBAL work list.
//   K1b narrowphase, THREAD per work item: libccd's MPR (ccdMPRPenetration) with
//       mpr_tolerance / mpr_iterations in the frame of body A, as a phase state machine (one support
//       evaluation per loop trip) so that the 32 pairs of a warp stay converged; mesh supports climb
//       the hull graph from the previous support vertex.
//   K1c compaction, WARP per env: valid staging slots -> contact list [env][DEX_MAXCON][8] + ncon, in
//       slot (= pair) order, which is MuJoCo's contact order.
#ifndef DEX_COLLIDE_CUH_
#define DEX_COLLIDE_CUH_

#include "dex_kin.cuh"
#include "dex_tree.cuh"

// per-warp shared-memory footprint of K1, in units of T
#define DEX_MAXCAND1 384   // capacity of the sphere-test survivor list (pass 1)
#define DEX_MAXGEOM 256    // world bounding spheres of all geoms are staged in shared memory
#define DEX_MAXBPL 128     // body pairs that survive the body-level culls
#define DEX_K1_SMEM_T (DEX_MAXBODY * 16 + DEX_MAXGEOM * 4 + DEX_MAXCAND1 + DEX_MAXBPL + 8 * 4)
// 20 single-warp CTAs per SM (the register limit) must fit the 196 KB shared-memory configuration, 1 KB of system use per CTA
// included: 64 words more per CTA select the 228 KB configuration, and the L1 lost to it costs K1a 15 % (0.156 -> 0.179 ms, r2zm)
static_assert((DEX_K1_SMEM_T * 4 + 1024) * 20 <= 196 * 1024, "K1a's shared memory crosses into the 228 KB configuration");
#define DEX_NCLASS 8       // solver-cost classes of K2's processing order
#define DEX_NCLASS_DENSE 4 // with the structured solver on (DevModel::arrow_ok): classes 0-3 dense kernel, 4-7 structured kernel
#define DEX_AR_MAXCON_K1 8 // = DEX_AR_MAXCON (dex_arrow.cuh)
#define DEX_AR_MAXCROSS_K1 2 // = DEX_AR_MAXCROSS
#define DEX_WORK_HDR 16     // words before the first work item: count, cursor, DEX_NCLASS class counters
#define DEX_MAXSLOT 64     // staging slots per env (one per convex-convex candidate, four per plane-box pair)
#define DEX_STAGE_ENV (DEX_MAXSLOT * DEX_CON_STRIDE)
#define DEX_GPACK 12  // geom_pack: centre(3) rbound | aabb centre(3) aabb half.x | aabb half.y half.z pad pad

template <typename T> struct V3 { T x, y, z; };
template <typename T> struct Xf { T R[9]; T t[3]; };  // pose of frame B expressed in frame A

// Everything the support function needs about one geom, resolved by the caller from the kernel's
// constant-bank model so that the non-inlined callees never take the model's address.
template <typename T>
struct GeomRef {
  const T* vert;   // body-frame hull vertices (mesh), 4 reals per vertex (x, y, z, pad), 16-byte aligned
  int n;           // vertex count
  int type;        // DEX_GEOM_*
  const T* mat;    // geom rotation in the body frame (box)
  const T* size;   // half sizes (box)
  const T* pack;   // geom_pack record (centre, rbound, aabb)
};
template <typename T>
DEX_DEV GeomRef<T> dex_geom_ref(const DevModel<T>& m, int g, int type) {
  GeomRef<T> r;
  r.vert = m.vert + 4 * dex_ldg(m.geom_vertadr + g);
  r.n = dex_ldg(m.geom_vertnum + g);
  r.type = type;
  r.mat = m.geom_mat + 9 * g;
  r.size = m.geom_size + 3 * g;
  r.pack = m.geom_pack + DEX_GPACK * g;
  return r;
}

// Body-frame support point for the body-frame unit direction (lx,ly,lz): mjccd_support.
// All 32 lanes call this together (warp shuffles inside).
template <typename T>
DEX_DEV_NOINLINE V3<T> dex_support_local(const T* vert, int n, int gtype, const T* mat, const T* size, const T* pack,
                                         T lx, T ly, T lz) {
  V3<T> out;
  if (gtype == DEX_GEOM_MESH) {
    T best = -Real<T>::big();
    int bi = 0x7fffffff;
    for (int i = dex_lane(); i < n; i += 32) {
      T vx, vy, vz;
      dex_ldg_xyz(vert + 4 * i, vx, vy, vz);
      const T t = vx * lx + vy * ly + vz * lz;
      if (t > best) { best = t; bi = i; }
    }
    dex_warp_argmax(best, bi);
    dex_ldg_xyz(vert + 4 * bi, out.x, out.y, out.z);
  } else if (gtype == DEX_GEOM_CAPSULE) {  // sphere of the radius plus the half length along the geom's z axis
    T G[9], ld[3] = {lx, ly, lz}, l2[3], cn[3], loc[3];
    for (int i = 0; i < 9; ++i) G[i] = dex_ldg(mat + i);
    rotT_vec(l2, G, ld);
    const T r = dex_ldg(size), hl = dex_ldg(size + 1);
    cn[0] = l2[0] * r; cn[1] = l2[1] * r; cn[2] = l2[2] * r + (l2[2] > 0 ? hl : -hl);
    rot_vec(loc, G, cn);
    out.x = loc[0] + dex_ldg(pack); out.y = loc[1] + dex_ldg(pack + 1); out.z = loc[2] + dex_ldg(pack + 2);
  } else {  // box
    T G[9], ld[3] = {lx, ly, lz}, l2[3], cn[3], loc[3];
    for (int i = 0; i < 9; ++i) G[i] = dex_ldg(mat + i);
    rotT_vec(l2, G, ld);
    cn[0] = l2[0] > 0 ? dex_ldg(size) : -dex_ldg(size);
    cn[1] = l2[1] > 0 ? dex_ldg(size + 1) : -dex_ldg(size + 1);
    cn[2] = l2[2] > 0 ? dex_ldg(size + 2) : -dex_ldg(size + 2);
    rot_vec(loc, G, cn);
    out.x = loc[0] + dex_ldg(pack); out.y = loc[1] + dex_ldg(pack + 1); out.z = loc[2] + dex_ldg(pack + 2);
  }
  return out;
}
template <typename T>
DEX_DEV V3<T> dex_support_ref(const GeomRef<T>& g, T lx, T ly, T lz) {
  return dex_support_local(g.vert, g.n, g.type, g.mat, g.size, g.pack, lx, ly, lz);
}

// one stored contact: dist, pos(3), normal(3), pair id
template <typename T>
DEX_DEV void dex_store_contact(T* dst, T dist, const T* pos, const T* normal, int pair) {
  dst[0] = dist; dst[1] = pos[0]; dst[2] = pos[1]; dst[3] = pos[2];
  dst[4] = normal[0]; dst[5] = normal[1]; dst[6] = normal[2];
  dst[7] = (T)pair;  // exact for pair ids < 2^24
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
// Minkowski-difference support point: v = v1 - v2, v1 on geom A, v2 on geom B (both in frame A)
template <typename T> struct SupPt { T v[3], v1[3]; };

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

// ------------------------------------------------------------------------------------------
// K1b: ccdMPRPenetration with ONE THREAD per candidate pair, written as a state machine so that the 32
// pairs of a warp stay converged: every trip of the loop evaluates one Minkowski support point (the
// common, expensive part) and then advances the pair's phase (portal construction, refinement,
// penetration search) -- libccd's discoverPortal / refinePortal / findPenetr unrolled into phases with
// identical arithmetic and identical exits.  The r01h profile showed why: with a group of lanes per
// pair the scalar portal logic is replicated G times and pairs of one warp diverge at once (12 of 32
// lanes active, 2 k warp instructions per pair).
// The mesh support function climbs the hull graph (as MuJoCo's mesh support does) from the previous
// support vertex of the same geom; hulls of up to DEX_SCAN_MAX vertices are scanned exhaustively.
// On a convex hull every local maximum is global, so both give the first-maximum vertex except on
// exact ties (a face perpendicular to the direction), the degenerate states DESIGN.md describes.
// ------------------------------------------------------------------------------------------
#ifndef DEX_SCAN_MAX
#define DEX_SCAN_MAX 4    // 0 / 4 / 8 / 16-32 / 48: narrowphase 0.109 / 0.109 / 0.110 / 0.115 / 0.120 ms per launch (climbing wins even on small hulls)
#endif
template <typename T>
struct LaneGeom {
  const T* vert;   // body-frame hull vertices of the geom instance (4 reals per vertex)
  int n, type, g;
  int adjbase;     // mesh-vertex base into DevModel::vadj
  int last;        // previous support vertex (warm start), -1 = none
};
template <typename T>
DEX_DEV void dex_support_lane(const DevModel<T>& m, LaneGeom<T>& G, T lx, T ly, T lz, T* out) {
  if (G.type == DEX_GEOM_MESH) {
    int bi = 0;
    if (G.n <= DEX_SCAN_MAX) {
      T best = -Real<T>::big();
      for (int i = 0; i < G.n; ++i) {
        T vx, vy, vz;
        dex_ldg_xyz(G.vert + 4 * i, vx, vy, vz);
        const T t = vx * lx + vy * ly + vz * lz;
        if (t > best) { best = t; bi = i; }
      }
    } else {
      // start: the table vertex of the direction's cube-map cell, or the previous support vertex if that is better
      T vx, vy, vz;
      {
        const T ax = dx_abs(lx), ay = dx_abs(ly), az = dx_abs(lz);
        int axis; T dm, du, dv;
        if (ax >= ay && ax >= az) { axis = 0; dm = lx; du = ly; dv = lz; }
        else if (ay >= az) { axis = 1; dm = ly; du = lz; dv = lx; }
        else { axis = 2; dm = lz; du = lx; dv = ly; }
        const T inv = T(4) / dx_max(dx_abs(dm), Real<T>::minval());
        int iu = (int)(du * inv + T(4)), iv = (int)(dv * inv + T(4));
        iu = iu < 0 ? 0 : (iu > 7 ? 7 : iu); iv = iv < 0 ? 0 : (iv > 7 ? 7 : iv);
        const int c = ((2 * axis + (dm < 0 ? 1 : 0)) << 6) | (iu << 3) | iv;
        bi = (dex_ldg(m.geom_cubemap + DEX_CUBEMAP_WORDS * G.g + (c >> 1)) >> (16 * (c & 1))) & 0xffff;
      }
      dex_ldg_xyz(G.vert + 4 * bi, vx, vy, vz);
      T best = vx * lx + vy * ly + vz * lz;
      if (G.last >= 0) {
        dex_ldg_xyz(G.vert + 4 * G.last, vx, vy, vz);
        const T t = vx * lx + vy * ly + vz * lz;
        const bool bt = t > best;
        best = bt ? t : best; bi = bt ? G.last : bi;
      }
      // climb: all neighbours of the current vertex per trip, eight at a time with the loads of a batch in
      // flight together (the dependent chain per move is adjacency word -> neighbour ids -> vertices); ties
      // move towards the lowest index
      for (int guard = 0; guard < G.n; ++guard) {
        const int info = dex_ldg(m.vadj + G.adjbase + bi);
        const int* nb = m.adj + (info >> 8);
        const int num = info & 255;
        int nxt = bi;
        T nv = best;
        for (int k0 = 0; k0 < num; k0 += 8) {
          int j[8];
          T t[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) j[u] = (k0 + u < num) ? dex_ldg(nb + k0 + u) : bi;   // filler = current vertex: never "better"
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            dex_ldg_xyz(G.vert + 4 * j[u], vx, vy, vz);
            t[u] = vx * lx + vy * ly + vz * lz;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const bool bt = (t[u] > nv) | ((t[u] == nv) & (j[u] < nxt));
            nv = bt ? t[u] : nv; nxt = bt ? j[u] : nxt;
          }
        }
        if (nxt == bi) break;
        bi = nxt; best = nv;
      }
      G.last = bi;
    }
    dex_ldg_xyz(G.vert + 4 * bi, out[0], out[1], out[2]);
  } else {  // box, capsule (mjccd_support: sphere of the radius plus the half length along the geom's z axis)
    const T* mat = m.geom_mat + 9 * G.g;
    const T* size = m.geom_size + 3 * G.g;
    const T* pack = m.geom_pack + DEX_GPACK * G.g;
    T Gm[9], ld[3] = {lx, ly, lz}, l2[3], cn[3], loc[3];
    for (int i = 0; i < 9; ++i) Gm[i] = dex_ldg(mat + i);
    rotT_vec(l2, Gm, ld);
    if (G.type == DEX_GEOM_CAPSULE) {
      const T r = dex_ldg(size), hl = dex_ldg(size + 1);
      cn[0] = l2[0] * r; cn[1] = l2[1] * r; cn[2] = l2[2] * r + (l2[2] > 0 ? hl : -hl);
    } else {
    cn[0] = l2[0] > 0 ? dex_ldg(size) : -dex_ldg(size);
    cn[1] = l2[1] > 0 ? dex_ldg(size + 1) : -dex_ldg(size + 1);
    cn[2] = l2[2] > 0 ? dex_ldg(size + 2) : -dex_ldg(size + 2);
    }
    rot_vec(loc, Gm, cn);
    out[0] = loc[0] + dex_ldg(pack); out[1] = loc[1] + dex_ldg(pack + 1); out[2] = loc[2] + dex_ldg(pack + 2);
  }
}

enum { MPR_P1 = 1, MPR_P2 = 2, MPR_DISCOVER = 3, MPR_REFINE = 4, MPR_PENETR = 5, MPR_DONE = 0 };

// One pair's MPR state.  begin() loads the item (env, slot << 24 | pair), step() evaluates one support
// point and advances the phase, finish() writes the env's staging slot once phase == MPR_DONE.  The
// kernel keeps one of these per thread and refills it from the global work list as soon as its pair is
// done, so that a warp's lanes stay busy although pairs need between 1 and ~50 support evaluations.
template <typename T>
struct MprLane {
  SupPt<T> s[4];
  T dir[3];
  Xf<T> rel;
  LaneGeom<T> A, B;
  T half_margin, depth, pdir[3], pos[3];
  int env, p, slot, b1, phase, guard;
  bool hit;

  DEX_DEV void begin(const DevModel<T>& m, const BatchState<T>& st, int env_, int wv) {
  env = env_; p = wv & 0xffffff; slot = wv >> 24;
  const int g1 = dex_ldg(m.pair_geom1 + p), g2 = dex_ldg(m.pair_geom2 + p);
  b1 = dex_ldg(m.geom_body + g1);
  const int b2 = dex_ldg(m.geom_body + g2);
  half_margin = T(0.5) * dex_ldg(m.pair_margin + p);
  const T* xk = st.xkin + (size_t)env * DEX_XKIN_ENV;
  {
    T x1[3], R1[9];
    T x2[3], R2[9];
    for (int k = 0; k < 3; ++k) { x1[k] = xk[b1 * DEX_XKIN_BODY + k]; x2[k] = xk[b2 * DEX_XKIN_BODY + k]; }
    for (int k = 0; k < 9; ++k) { R1[k] = xk[b1 * DEX_XKIN_BODY + 7 + k]; R2[k] = xk[b2 * DEX_XKIN_BODY + 7 + k]; }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
      rel.R[3 * r + c] = R1[r] * R2[c] + R1[3 + r] * R2[3 + c] + R1[6 + r] * R2[6 + c];
    const T d[3] = {x2[0] - x1[0], x2[1] - x1[1], x2[2] - x1[2]};
    rotT_vec(rel.t, R1, d);
  }
  A.vert = m.vert + 4 * dex_ldg(m.geom_vertadr + g1); A.n = dex_ldg(m.geom_vertnum + g1); A.type = dex_ldg(m.geom_type + g1);
  A.g = g1; A.adjbase = dex_ldg(m.geom_adjbase + g1); A.last = -1;
  B.vert = m.vert + 4 * dex_ldg(m.geom_vertadr + g2); B.n = dex_ldg(m.geom_vertnum + g2); B.type = dex_ldg(m.geom_type + g2);
  B.g = g2; B.adjbase = dex_ldg(m.geom_adjbase + g2); B.last = -1;
  depth = 0; pdir[0] = pdir[1] = pdir[2] = 0; pos[0] = pos[1] = pos[2] = 0;
  hit = false;
  {
    const T* pa = m.geom_pack + DEX_GPACK * g1;
    const T* pb = m.geom_pack + DEX_GPACK * g2;
    T cb[3] = {dex_ldg(pb), dex_ldg(pb + 1), dex_ldg(pb + 2)}, cbw[3];
    rot_vec(cbw, rel.R, cb);
    for (int k = 0; k < 3; ++k) { s[0].v1[k] = dex_ldg(pa + k); s[0].v[k] = s[0].v1[k] - (cbw[k] + rel.t[k]); }
  }
  if (ccd_zero(s[0].v[0]) && ccd_zero(s[0].v[1]) && ccd_zero(s[0].v[2])) s[0].v[0] += T(2.220446049250313e-15);
  dir[0] = -s[0].v[0]; dir[1] = -s[0].v[1]; dir[2] = -s[0].v[2];
  normalize3(dir);
  phase = MPR_P1; guard = 0;
  }

  // one support evaluation + phase update; `continue` of the loop form = return
  DEX_DEV void step(const DevModel<T>& m) {
    const T tolerance = m.mpr_tolerance;
    const int iterations = m.mpr_iterations;
    SupPt<T> v4;
    T va[3], vb[3];
    // ---- Minkowski-difference support point along dir (v = v1 - v2, both shapes inflated by margin / 2)
    {
      T a[3], b[3], nd[3] = {-dir[0], -dir[1], -dir[2]}, lb[3], bw[3];
      dex_support_lane(m, A, dir[0], dir[1], dir[2], a);
      rotT_vec(lb, rel.R, nd);
      dex_support_lane(m, B, lb[0], lb[1], lb[2], b);
      rot_vec(bw, rel.R, b);
      for (int k = 0; k < 3; ++k) {
        v4.v1[k] = a[k] + dir[k] * half_margin;
        v4.v[k] = v4.v1[k] - (bw[k] + rel.t[k] - dir[k] * half_margin);
      }
    }
    const T dv = dot3(v4.v, dir);
    bool need_portal = false;
    if (phase == MPR_P1) {
      s[1] = v4;
      if (ccd_zero(dv) || dv < 0) { phase = MPR_DONE; return; }
      cross3(dir, s[0].v, s[1].v);
      if (ccd_zero(dot3(dir, dir))) {
        if (ccd_zero(s[1].v[0]) && ccd_zero(s[1].v[1]) && ccd_zero(s[1].v[2])) {  // touching: findPenetrTouch
          depth = 0; pdir[0] = pdir[1] = pdir[2] = 0;
        } else {  // origin on v0-v1: findPenetrSegment
          depth = dx_sqrt(dot3(s[1].v, s[1].v));
          pdir[0] = s[1].v[0]; pdir[1] = s[1].v[1]; pdir[2] = s[1].v[2];
          normalize3(pdir);
        }
        for (int k = 0; k < 3; ++k) pos[k] = s[1].v1[k] - T(0.5) * s[1].v[k];  // (v1 + v2) / 2
        hit = true; phase = MPR_DONE; return;
      }
      normalize3(dir);
      phase = MPR_P2;
    } else if (phase == MPR_P2) {
      s[2] = v4;
      if (ccd_zero(dv) || dv < 0) { phase = MPR_DONE; return; }
      for (int k = 0; k < 3; ++k) { va[k] = s[1].v[k] - s[0].v[k]; vb[k] = s[2].v[k] - s[0].v[k]; }
      cross3(dir, va, vb);
      normalize3(dir);
      if (dot3(dir, s[0].v) > 0) {
        SupPt<T> tmp = s[1]; s[1] = s[2]; s[2] = tmp;
        dir[0] = -dir[0]; dir[1] = -dir[1]; dir[2] = -dir[2];
      }
      phase = MPR_DISCOVER; guard = 0;
    } else if (phase == MPR_DISCOVER) {  // discoverPortal: find v3
      s[3] = v4;
      if (ccd_zero(dv) || dv < 0) { phase = MPR_DONE; return; }
      bool cont = false;
      cross3(va, s[1].v, s[3].v);
      T dt = dot3(va, s[0].v);
      if (dt < 0 && !ccd_zero(dt)) { s[2] = s[3]; cont = true; }
      if (!cont) {
        cross3(va, s[3].v, s[2].v);
        dt = dot3(va, s[0].v);
        if (dt < 0 && !ccd_zero(dt)) { s[1] = s[3]; cont = true; }
      }
      if (cont) {
        if (++guard > 100) { phase = MPR_DONE; return; }
        for (int k = 0; k < 3; ++k) { va[k] = s[1].v[k] - s[0].v[k]; vb[k] = s[2].v[k] - s[0].v[k]; }
        cross3(dir, va, vb);
        normalize3(dir);
      } else {
        phase = MPR_REFINE; guard = 0;
        need_portal = true;
      }
    } else {  // refinePortal / findPenetr: both test the new point against the portal, then expand it
      const bool reach = portal_reach_tol(s, v4, dir, tolerance);
      if (phase == MPR_REFINE) {
        if (!(ccd_zero(dv) || dv > 0) || reach) { phase = MPR_DONE; return; }
      } else if (reach || guard > iterations) {
        T d2 = point_tri_dist2(s[1].v, s[2].v, s[3].v, pdir);
        depth = dx_sqrt(d2);
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
          // p1 = sum b_i v1_i, p2 = sum b_i v2_i with v2 = v1 - v;  pos = (p1 + p2) / 2
          T p1 = b[0] * s[0].v1[k] + b[1] * s[1].v1[k] + b[2] * s[2].v1[k] + b[3] * s[3].v1[k];
          T pv = b[0] * s[0].v[k] + b[1] * s[1].v[k] + b[2] * s[2].v[k] + b[3] * s[3].v[k];
          pos[k] = (p1 - T(0.5) * pv) * inv;
        }
        hit = true; phase = MPR_DONE; return;
      }
      expand_portal(s, v4);
      ++guard;
      if (phase == MPR_REFINE && guard > 100) { phase = MPR_DONE; return; }
      need_portal = true;
    }
    if (need_portal) {
      portal_dir(s, dir);
      if (phase == MPR_REFINE) {
        const T dt = dot3(dir, s[1].v);
        if (ccd_zero(dt) || dt > 0) { phase = MPR_PENETR; guard = 0; }
      }
    }
  }

  DEX_DEV void finish(const DevModel<T>& m, const BatchState<T>& st) const {
  T* out = st.stage + (size_t)env * DEX_STAGE_ENV + slot * DEX_CON_STRIDE;
  if (hit && !(pdir[0] == 0 && pdir[1] == 0 && pdir[2] == 0)) {
    const T* xk = st.xkin + (size_t)env * DEX_XKIN_ENV + b1 * DEX_XKIN_BODY;
    T x1[3], R1[9];
    for (int k = 0; k < 3; ++k) x1[k] = xk[k];
    for (int k = 0; k < 9; ++k) R1[k] = xk[7 + k];
    const T margin = 2 * half_margin;
    T pw[3], nw[3];
    rot_vec(pw, R1, pos);
    rot_vec(nw, R1, pdir);
    pw[0] += x1[0]; pw[1] += x1[1]; pw[2] += x1[2];
    dex_store_contact(out, margin - depth, pw, nw, p);
  } else out[7] = T(-1);
  }
};

// plain per-item form (test emulator)
template <typename T>
DEX_DEV void dex_narrow_lane(const DevModel<T>& m, const BatchState<T>& st, int env, int wv) {
  MprLane<T> L;
  L.begin(m, st, env, wv);
  while (L.phase != MPR_DONE) L.step(m);
  L.finish(m, st);
}

// pose of body b2 in the frame of body b1
template <typename T>
DEX_DEV void dex_relative(const T* xpos, const T* xmat, int b1, int b2, Xf<T>& x) {
  const T* R1 = xmat + 9 * b1;
  const T* R2 = xmat + 9 * b2;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
    x.R[3 * r + c] = R1[r] * R2[c] + R1[3 + r] * R2[3 + c] + R1[6 + r] * R2[6 + c];
  T d[3] = {xpos[3 * b2] - xpos[3 * b1], xpos[3 * b2 + 1] - xpos[3 * b1 + 1], xpos[3 * b2 + 2] - xpos[3 * b1 + 2]};
  rotT_vec(x.t, R1, d);
}

template <typename T>
DEX_DEV bool dex_sphere_aabb(const T* c, T r, const T* pack) {  // pack + 4: aabb centre(3), half(3)
  T s = 0;
  for (int k = 0; k < 3; ++k) {
    T e = dx_abs(c[k] - dex_ldg(pack + 4 + k)) - dex_ldg(pack + 7 + k);
    if (e > 0) s += e * e;
  }
  return s <= r * r;
}

// Separating-axis test between the body-frame AABB of geom 1 (frame A) and that of geom 2 (frame B,
// pose `rel` in A), both inflated by `margin`/2.  Conservative (the boxes enclose the hulls), so a
// "separated" verdict can never remove a real contact.  Gottschalk's 15-axis OBB test.
template <typename T>
DEX_DEV bool dex_obb_overlap(const Xf<T>& rel, const T* pk1, const T* pk2, T margin) {
  T ha[3], hb[3], cb[3], t[3], AR[9];
  for (int k = 0; k < 3; ++k) {
    ha[k] = dex_ldg(pk1 + 7 + k) + T(0.5) * margin;
    hb[k] = dex_ldg(pk2 + 7 + k) + T(0.5) * margin;
    cb[k] = dex_ldg(pk2 + 4 + k);
  }
  rot_vec(t, rel.R, cb);
  for (int k = 0; k < 3; ++k) t[k] += rel.t[k] - dex_ldg(pk1 + 4 + k);
  const T eps = T(1e-6);
  for (int k = 0; k < 9; ++k) AR[k] = dx_abs(rel.R[k]) + eps;
  const T* R = rel.R;
  for (int i = 0; i < 3; ++i)
    if (dx_abs(t[i]) > ha[i] + hb[0] * AR[3 * i] + hb[1] * AR[3 * i + 1] + hb[2] * AR[3 * i + 2]) return false;
  for (int j = 0; j < 3; ++j)
    if (dx_abs(t[0] * R[j] + t[1] * R[3 + j] + t[2] * R[6 + j]) > ha[0] * AR[j] + ha[1] * AR[3 + j] + ha[2] * AR[6 + j] + hb[j]) return false;
  for (int i = 0; i < 3; ++i) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    for (int j = 0; j < 3; ++j) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const T ra = ha[i1] * AR[3 * i2 + j] + ha[i2] * AR[3 * i1 + j];
      const T rb = hb[j1] * AR[3 * i + j2] + hb[j2] * AR[3 * i + j1];
      if (dx_abs(t[i2] * R[3 * i1 + j] - t[i1] * R[3 * i2 + j]) > ra + rb) return false;
    }
  }
  return true;
}

// mjraw_SphereSphere: contact between two spheres (centres in the world frame); returns whether it exists
template <typename T>
DEX_DEV bool dex_sphere_sphere(const T* p1, T r1, const T* p2, T r2, T margin, int pair, T* out) {
  T dif[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  const T cdist = dx_sqrt(dot3(dif, dif));
  const T dist = cdist - r1 - r2;
  if (dist > margin) return false;
  T n[3] = {1, 0, 0};
  if (cdist >= Real<T>::minval()) { n[0] = dif[0] / cdist; n[1] = dif[1] / cdist; n[2] = dif[2] / cdist; }
  const T a = r1 + T(0.5) * dist;
  T pos[3] = {p1[0] + n[0] * a, p1[1] + n[1] * a, p1[2] + n[2] * a};
  dex_store_contact(out, dist, pos, n, pair);
  return true;
}
// mjc_CapsuleCapsule (engine_collision_primitive.c): closest points of the two axis segments, then sphere-sphere there; for
// parallel axes the four end-point configurations, two contacts at most.  Warp-uniform arithmetic, lane 0 writes the two
// staging slots at `out` (pair < 0 = no contact).
// Not inlined (it is off the reorient scenes' path and K1a is sensitive to its code size); the model tables come as plain
// pointers so that the callee never takes the address of the kernel's constant-bank model.
template <typename T>
DEX_DEV_NOINLINE void dex_capsule_capsule(const int* geom_body, const T* geom_mat, const T* geom_pack, const T* geom_size, const T* xpos,
                                          const T* xmat, int g1, int g2, T margin, int pair, T* out) {
  const int b1 = dex_ldg(geom_body + g1), b2 = dex_ldg(geom_body + g2);
  T pos1[3], pos2[3], ax1[3], ax2[3];
  {
    const T* G1 = geom_mat + 9 * g1; const T* G2 = geom_mat + 9 * g2;
    T l1[3] = {dex_ldg(geom_pack + DEX_GPACK * g1), dex_ldg(geom_pack + DEX_GPACK * g1 + 1), dex_ldg(geom_pack + DEX_GPACK * g1 + 2)};
    T l2[3] = {dex_ldg(geom_pack + DEX_GPACK * g2), dex_ldg(geom_pack + DEX_GPACK * g2 + 1), dex_ldg(geom_pack + DEX_GPACK * g2 + 2)};
    T z1[3] = {dex_ldg(G1 + 2), dex_ldg(G1 + 5), dex_ldg(G1 + 8)}, z2[3] = {dex_ldg(G2 + 2), dex_ldg(G2 + 5), dex_ldg(G2 + 8)};
    rot_vec(pos1, xmat + 9 * b1, l1); rot_vec(pos2, xmat + 9 * b2, l2);
    for (int k = 0; k < 3; ++k) { pos1[k] += xpos[3 * b1 + k]; pos2[k] += xpos[3 * b2 + k]; }
    rot_vec(ax1, xmat + 9 * b1, z1); rot_vec(ax2, xmat + 9 * b2, z2);
  }
  const T r1 = dex_ldg(geom_size + 3 * g1), h1 = dex_ldg(geom_size + 3 * g1 + 1);
  const T r2 = dex_ldg(geom_size + 3 * g2), h2 = dex_ldg(geom_size + 3 * g2 + 1);
  const T dif[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
  const T ma = dot3(ax1, ax1), mb = -dot3(ax1, ax2), mc = dot3(ax2, ax2);
  const T u = -dot3(ax1, dif), v = dot3(ax2, dif);
  const T det = ma * mc - mb * mb;
  T rec[2][DEX_CON_STRIDE];
  int n = 0;
  auto at = [&](T x1, T x2) {
    T v1[3] = {pos1[0] + ax1[0] * x1, pos1[1] + ax1[1] * x1, pos1[2] + ax1[2] * x1};
    T v2[3] = {pos2[0] + ax2[0] * x2, pos2[1] + ax2[1] * x2, pos2[2] + ax2[2] * x2};
    if (n < 2 && dex_sphere_sphere(v1, r1, v2, r2, margin, pair, rec[n])) ++n;
  };
  if (dx_abs(det) >= Real<T>::minval()) {   // general configuration
    T x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
    if (x1 > h1) { x1 = h1; x2 = (v - mb * h1) / mc; }
    else if (x1 < -h1) { x1 = -h1; x2 = (v + mb * h1) / mc; }
    if (x2 > h2) { x2 = h2; x1 = dx_clamp((u - mb * h2) / ma, -h1, h1); }
    else if (x2 < -h2) { x2 = -h2; x1 = dx_clamp((u + mb * h2) / ma, -h1, h1); }
    at(x1, x2);
  } else {   // parallel axes: each end point of either segment against the other segment
    at(h1, dx_clamp((v - mb * h1) / mc, -h2, h2));
    at(-h1, dx_clamp((v + mb * h1) / mc, -h2, h2));
    if (n < 2) at(dx_clamp((u - mb * h2) / ma, -h1, h1), h2);
    if (n < 2) at(dx_clamp((u + mb * h2) / ma, -h1, h1), -h2);
  }
  if (dex_lane() == 0) {
    for (int i = 0; i < 2; ++i) {
      if (i < n) { for (int k = 0; k < DEX_CON_STRIDE; ++k) out[i * DEX_CON_STRIDE + k] = rec[i][k]; }
      else out[i * DEX_CON_STRIDE + 7] = T(-1);
    }
  }
}

// ------------------------------------------------------------------------------------------
// K1a body.  smem: [xpos | xmat | geom spheres | body spheres | cand(int) | surviving body pairs(int)]
// ------------------------------------------------------------------------------------------

template <typename T>
DEX_DEV void dex_collide_env(const DevModel<T>& m, const BatchState<T>& st, int env, T* smem) {
  if (env >= st.nenv || dex_env_off(st, env)) return;  // warp-uniform
  const int lane = dex_lane();
  T* xpos = smem;
  T* xmat = xpos + DEX_MAXBODY * 3;
  T* gsph = xmat + DEX_MAXBODY * 9;                 // [DEX_MAXGEOM][4] world bounding spheres: centre, radius
  T* bsph = gsph + DEX_MAXGEOM * 4;                 // [DEX_MAXBODY][4] world body bounding spheres
  int* cand = (int*)(bsph + DEX_MAXBODY * 4);       // DEX_MAXCAND1 ints (fits: sizeof(int) <= sizeof(T))
  int* bpl = cand + DEX_MAXCAND1;                   // DEX_MAXBPL surviving body pairs
  {  // body poses of K0 (dex_tree.cuh): records of 16 = xpos(3) xquat(4) xmat(9), fetched as quarter records of 4
     // with every load of the warp in flight before the first use (the r01m profile had 22 % of K1a's stall
     // samples on this copy when it was a load-store loop)
    const T* xk = st.xkin + (size_t)env * DEX_XKIN_ENV;
    const int nq4 = m.nbody * 4;
    T v[DEX_MAXBODY * 4 / 32][4];
#pragma unroll
    for (int k = 0; k < DEX_MAXBODY * 4 / 32; ++k) {
      const int i = lane + 32 * k;
      if (i < nq4) dex_ld4(xk + 4 * i, v[k]);
    }
#pragma unroll
    for (int k = 0; k < DEX_MAXBODY * 4 / 32; ++k) {
      const int i = lane + 32 * k;
      if (i < nq4) {
        const int b = i >> 2, q = i & 3;
        if (q == 0) { xpos[3 * b] = v[k][0]; xpos[3 * b + 1] = v[k][1]; xpos[3 * b + 2] = v[k][2]; }
        else if (q == 1) xmat[9 * b] = v[k][3];
        else if (q == 2) { xmat[9 * b + 1] = v[k][0]; xmat[9 * b + 2] = v[k][1]; xmat[9 * b + 3] = v[k][2]; xmat[9 * b + 4] = v[k][3]; }
        else { xmat[9 * b + 5] = v[k][0]; xmat[9 * b + 6] = v[k][1]; xmat[9 * b + 7] = v[k][2]; xmat[9 * b + 8] = v[k][3]; }
      }
    }
  }
  dex_syncwarp();
  // ---- world-frame bounding spheres of every geom and every body (one lane each)
  for (int g = lane; g < m.ngeom; g += 32) {
    const int b = dex_ldg(m.geom_body + g);
    const T* pk = m.geom_pack + DEX_GPACK * g;
    T l[3] = {dex_ldg(pk), dex_ldg(pk + 1), dex_ldg(pk + 2)}, c[3];
    rot_vec(c, xmat + 9 * b, l);
    gsph[4 * g] = c[0] + xpos[3 * b]; gsph[4 * g + 1] = c[1] + xpos[3 * b + 1]; gsph[4 * g + 2] = c[2] + xpos[3 * b + 2];
    gsph[4 * g + 3] = dex_ldg(pk + 3);
  }
  for (int b = lane; b < m.nbody; b += 32) {
    const T* bs = m.body_bsphere + 4 * b;
    T l[3] = {dex_ldg(bs), dex_ldg(bs + 1), dex_ldg(bs + 2)}, c[3];
    rot_vec(c, xmat + 9 * b, l);
    bsph[4 * b] = c[0] + xpos[3 * b]; bsph[4 * b + 1] = c[1] + xpos[3 * b + 1]; bsph[4 * b + 2] = c[2] + xpos[3 * b + 2];
    bsph[4 * b + 3] = dex_ldg(bs + 3);
  }
  dex_syncwarp();

  // ---- broadphase, body pairs: bounding spheres (plane bodies: half-space test) for all pairs, survivors
  //      compacted, then the bodies' AABBs as oriented boxes on dense lanes (r01d: 273 -> 61 -> 13 pairs)
  const T slack = T(1e-6);   // world-frame centres carry FP32 rounding: keep every cull conservative
  int nbp = 0;
  // three chunks of 32 body pairs per trip, their body ids requested up front (same reason as in the geom-pair pass below)
  for (int bp0 = 0; bp0 < m.nbodypair; bp0 += 96) {
    int b1s[3], b2s[3];
    bool ps[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int bp = bp0 + 32 * q + lane;
      b1s[q] = b2s[q] = 0;
      if (bp < m.nbodypair) { b1s[q] = dex_ldg(m.bp_body1 + bp); b2s[q] = dex_ldg(m.bp_body2 + bp); }
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int bp = bp0 + 32 * q + lane;
      bool pass = false;
      if (bp < m.nbodypair) {
        const int b1 = b1s[q], b2 = b2s[q];
        const T r1 = bsph[4 * b1 + 3], r2 = bsph[4 * b2 + 3];
        if (r1 < 0) {  // body 1 owns a plane (planes sort first in a pair): half-space test of body 2's sphere
          const int gp = dex_ldg(m.pair_geom1 + dex_ldg(m.bp_pairadr + bp));
          const T* G = m.geom_mat + 9 * gp;
          T nl[3] = {dex_ldg(G + 2), dex_ldg(G + 5), dex_ldg(G + 8)}, n[3];
          rot_vec(n, xmat + 9 * b1, nl);
          T d[3] = {bsph[4 * b2] - gsph[4 * gp], bsph[4 * b2 + 1] - gsph[4 * gp + 1], bsph[4 * b2 + 2] - gsph[4 * gp + 2]};
          pass = dot3(d, n) <= r2 + slack;
        } else {
          T d[3] = {bsph[4 * b2] - bsph[4 * b1], bsph[4 * b2 + 1] - bsph[4 * b1 + 1], bsph[4 * b2 + 2] - bsph[4 * b1 + 2]};
          const T bound = r1 + r2 + slack;
          pass = dot3(d, d) <= bound * bound;
        }
      }
      ps[q] = pass;
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const unsigned mk = dex_ballot(ps[q]);
      if (ps[q]) cand[nbp + dex_popc(mk & ((1u << lane) - 1u))] = bp0 + 32 * q + lane;   // nbodypair <= DEX_MAXCAND1 (host check)
      nbp += dex_popc(mk);
    }
  }
  dex_syncwarp();
  int nbpl = 0;
  unsigned status = 0;
  for (int c0 = 0; c0 < nbp; c0 += 32) {
    bool pass = false;
    int bp = 0;
    if (c0 + lane < nbp) {
      bp = cand[c0 + lane];
      const int b1 = dex_ldg(m.bp_body1 + bp), b2 = dex_ldg(m.bp_body2 + bp);
      pass = true;
      if (bsph[4 * b1 + 3] >= 0) {
        Xf<T> rb;
        dex_relative(xpos, xmat, b1, b2, rb);
        pass = dex_obb_overlap(rb, m.body_pack + DEX_GPACK * b1, m.body_pack + DEX_GPACK * b2, m.max_margin + 2 * slack);
      }
    }
    const unsigned mk = dex_ballot(pass);
    if (pass) {
      const int idx = nbpl + dex_popc(mk & ((1u << lane) - 1u));
      if (idx < DEX_MAXBPL) bpl[idx] = bp;
    }
    nbpl += dex_popc(mk);
  }
  if (nbpl > DEX_MAXBPL) { status |= DEX_STATUS_CAND_OVERFLOW; nbpl = DEX_MAXBPL; }
  dex_syncwarp();

  // ---- broadphase, geom pairs, pass 1: world-frame bounding spheres of the geoms of every surviving body pair
  //      (cheap, mostly rejects); survivors are compacted so that the costlier AABB / OBB tests of pass 2 run on
  //      dense lanes (r01d: 900 pairs visited per env, 86 survive the spheres, 7 survive everything)
  int ncand = 0;
  int adr_n = 0, num_n = 0, b1_n = 0;   // header of the next body pair, requested one trip ahead
  if (nbpl > 0) { const int k = bpl[0]; adr_n = dex_ldg(m.bp_pairadr + k); num_n = dex_ldg(m.bp_pairnum + k); b1_n = dex_ldg(m.bp_body1 + k); }
  for (int bi = 0; bi < nbpl; ++bi) {
    const int adr = adr_n, num = num_n, b1 = b1_n;
    if (bi + 1 < nbpl) { const int k = bpl[bi + 1]; adr_n = dex_ldg(m.bp_pairadr + k); num_n = dex_ldg(m.bp_pairnum + k); b1_n = dex_ldg(m.bp_body1 + k); }
    const bool plane = bsph[4 * b1 + 3] < 0;
    // DEX_K1_NCH chunks of 32 pairs per trip: all chunks' pair words are requested before any is used, so the dependent
    // global load -> shared-memory lookups of one chunk overlap those of the others (1 / 2 / 3 / 4 chunks: 0.158 / 0.156 / 0.156 /
    // 0.159 ms per launch)
#ifndef DEX_K1_NCH
#define DEX_K1_NCH 2
#endif
    for (int p0 = 0; p0 < num; p0 += 32 * DEX_K1_NCH) {
      int gg[DEX_K1_NCH];
      T mg[DEX_K1_NCH];
      bool ok[DEX_K1_NCH];
#pragma unroll
      for (int q = 0; q < DEX_K1_NCH; ++q) {
        gg[q] = 0; mg[q] = 0;
        if (p0 + 32 * q + lane < num) { gg[q] = dex_ldg(m.pair_gg + adr + p0 + 32 * q + lane); mg[q] = dex_ldg(m.pair_margin + adr + p0 + 32 * q + lane); }
      }
#pragma unroll
      for (int q = 0; q < DEX_K1_NCH; ++q) {
        ok[q] = false;
        if (p0 + 32 * q + lane < num) {
          const int g1 = gg[q] & 0xffff, g2 = gg[q] >> 16;
          T d[3] = {gsph[4 * g2] - gsph[4 * g1], gsph[4 * g2 + 1] - gsph[4 * g1 + 1], gsph[4 * g2 + 2] - gsph[4 * g1 + 2]};
          const T r2 = gsph[4 * g2 + 3];
          if (plane && dex_ldg(m.geom_type + g1) == DEX_GEOM_PLANE) {
            const T* G = m.geom_mat + 9 * g1;
            T nl[3] = {dex_ldg(G + 2), dex_ldg(G + 5), dex_ldg(G + 8)}, n[3];
            rot_vec(n, xmat + 9 * b1, nl);
            ok[q] = dot3(d, n) <= mg[q] + r2 + slack;
          } else {
            const T bound = mg[q] + gsph[4 * g1 + 3] + r2 + slack;
            ok[q] = dot3(d, d) <= bound * bound;
          }
        }
      }
#pragma unroll
      for (int q = 0; q < DEX_K1_NCH; ++q) {
        const unsigned mk = dex_ballot(ok[q]);
        if (ok[q]) {
          const int idx = ncand + dex_popc(mk & ((1u << lane) - 1u));
          if (idx < DEX_MAXCAND1) cand[idx] = adr + p0 + 32 * q + lane;
        }
        ncand += dex_popc(mk);
      }
    }
  }
  if (ncand > DEX_MAXCAND1) { status |= DEX_STATUS_CAND_OVERFLOW; ncand = DEX_MAXCAND1; }
  dex_syncwarp();
  // pass 2: sphere-vs-AABB both ways, then AABB-vs-AABB as oriented boxes; in-place compaction
  {
    int nkeep = 0;
    for (int c0 = 0; c0 < ncand; c0 += 32) {
      bool ok = false;
      int p = 0;
      if (c0 + lane < ncand) {
        p = cand[c0 + lane];
        const int g1 = dex_ldg(m.pair_geom1 + p), g2 = dex_ldg(m.pair_geom2 + p);
        ok = true;
        if (dex_ldg(m.geom_type + g1) != DEX_GEOM_PLANE) {
          const T margin = dex_ldg(m.pair_margin + p);
          const T* pk1 = m.geom_pack + DEX_GPACK * g1;
          const T* pk2 = m.geom_pack + DEX_GPACK * g2;
          Xf<T> rel;
          dex_relative(xpos, xmat, dex_ldg(m.geom_body + g1), dex_ldg(m.geom_body + g2), rel);
          T c2l[3] = {dex_ldg(pk2), dex_ldg(pk2 + 1), dex_ldg(pk2 + 2)}, c2[3];
          rot_vec(c2, rel.R, c2l);
          c2[0] += rel.t[0]; c2[1] += rel.t[1]; c2[2] += rel.t[2];
          const T r1 = dex_ldg(pk1 + 3), r2 = dex_ldg(pk2 + 3);
          ok = dex_sphere_aabb(c2, r2 + margin, pk1);                          // sphere 2 vs AABB 1 (frame 1)
          if (ok) {
            T e[3] = {dex_ldg(pk1) - rel.t[0], dex_ldg(pk1 + 1) - rel.t[1], dex_ldg(pk1 + 2) - rel.t[2]}, c1in2[3];
            rotT_vec(c1in2, rel.R, e);
            ok = dex_sphere_aabb(c1in2, r1 + margin, pk2);                     // sphere 1 vs AABB 2 (frame 2)
          }
          if (ok) ok = dex_obb_overlap(rel, pk1, pk2, margin);                 // AABB 1 vs AABB 2 as OBBs
        }
      }
      const unsigned m2 = dex_ballot(ok);
      dex_syncwarp();   // every lane has read its entry of this round before anything is overwritten
      if (ok) {
        const int idx = nkeep + dex_popc(m2 & ((1u << lane) - 1u));
        cand[idx] = p;                             // idx <= c0 + lane: never ahead of the read cursor
      }
      nkeep += dex_popc(m2);
      dex_syncwarp();
    }
    ncand = nkeep;
  }
  dex_syncwarp();

  // ---- plane pairs are finished here; convex-convex survivors become work items for K1b
  T* stage = st.stage + (size_t)env * DEX_STAGE_ENV;
  int nslot = 0;   // staging slots handed out so far (warp-uniform)
  int nwork = 0;   // convex-convex work items (warp-uniform)
  int* work = cand;  // reuse: work[i] = slot << 24 | pair is written behind the read cursor (i <= ci)
  for (int ci = 0; ci < ncand; ++ci) {
    const int p = cand[ci];
    const int g1 = dex_ldg(m.pair_geom1 + p), g2 = dex_ldg(m.pair_geom2 + p);
    const int t1 = dex_ldg(m.geom_type + g1), t2 = dex_ldg(m.geom_type + g2);
    if (t1 == DEX_GEOM_CAPSULE && t2 == DEX_GEOM_CAPSULE) {   // mjc_CapsuleCapsule: closed form, up to two contacts (parallel axes)
      if (nslot + 2 > DEX_MAXSLOT) { status |= DEX_STATUS_CAND_OVERFLOW; continue; }
      dex_capsule_capsule<T>(m.geom_body, m.geom_mat, m.geom_pack, m.geom_size, xpos, xmat, g1, g2, dex_ldg(m.pair_margin + p), p,
                             stage + nslot * DEX_CON_STRIDE);
      nslot += 2;
      continue;
    }
    if (t1 != DEX_GEOM_PLANE) {
      if (nslot < DEX_MAXSLOT) {
        dex_syncwarp();
        if (lane == 0) work[nwork] = (nslot << 24) | p;
        nwork++; nslot++;
      } else status |= DEX_STATUS_CAND_OVERFLOW;
      continue;
    }
    const int b1 = dex_ldg(m.geom_body + g1), b2 = dex_ldg(m.geom_body + g2);
    const T margin = dex_ldg(m.pair_margin + p);
    const T* G1 = m.geom_mat + 9 * g1;
    T nl[3] = {dex_ldg(G1 + 2), dex_ldg(G1 + 5), dex_ldg(G1 + 8)}, n[3], p1[3];
    rot_vec(n, xmat + 9 * b1, nl);
    T pl[3] = {dex_ldg(m.geom_pack + DEX_GPACK * g1), dex_ldg(m.geom_pack + DEX_GPACK * g1 + 1), dex_ldg(m.geom_pack + DEX_GPACK * g1 + 2)};
    rot_vec(p1, xmat + 9 * b1, pl);
    p1[0] += xpos[3 * b1]; p1[1] += xpos[3 * b1 + 1]; p1[2] += xpos[3 * b1 + 2];
    if (t2 == DEX_GEOM_BOX) {  // mjc_PlaneBox: lanes 0..7 test the corners, first 4 hits kept (4 slots)
      if (nslot + 4 > DEX_MAXSLOT) { status |= DEX_STATUS_CAND_OVERFLOW; continue; }
      T cl[3] = {dex_ldg(m.geom_pack + DEX_GPACK * g2), dex_ldg(m.geom_pack + DEX_GPACK * g2 + 1), dex_ldg(m.geom_pack + DEX_GPACK * g2 + 2)}, c2[3];
      rot_vec(c2, xmat + 9 * b2, cl);
      c2[0] += xpos[3 * b2]; c2[1] += xpos[3 * b2 + 1]; c2[2] += xpos[3 * b2 + 2];
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
        T cd = dist + ldist;
        T cp[3] = {corner[0] + c2[0] - n[0] * cd * T(0.5), corner[1] + c2[1] - n[1] * cd * T(0.5),
                   corner[2] + c2[2] - n[2] * cd * T(0.5)};
        dex_store_contact(stage + (nslot + rank) * DEX_CON_STRIDE, cd, cp, n, p);
      }
      int add = dex_popc(hm); if (add > 4) add = 4;
      if (lane >= add && lane < 4) stage[(nslot + lane) * DEX_CON_STRIDE + 7] = T(-1);  // unused slots of the four
      nslot += 4;
    } else if (t2 == DEX_GEOM_MESH) {  // mjc_PlaneConvex: deepest vertex (hull-neighbour extras: next round)
      if (nslot >= DEX_MAXSLOT) { status |= DEX_STATUS_CAND_OVERFLOW; continue; }
      T nn[3] = {-n[0], -n[1], -n[2]}, ld[3], w[3];
      rotT_vec(ld, xmat + 9 * b2, nn);
      const V3<T> sp = dex_support_ref(dex_geom_ref(m, g2, t2), ld[0], ld[1], ld[2]);
      T sl[3] = {sp.x, sp.y, sp.z};
      rot_vec(w, xmat + 9 * b2, sl);
      w[0] += xpos[3 * b2]; w[1] += xpos[3 * b2 + 1]; w[2] += xpos[3 * b2 + 2];
      T dif[3] = {w[0] - p1[0], w[1] - p1[1], w[2] - p1[2]};
      const T dist = dot3(dif, n);
      if (lane == 0) {
        if (dist <= margin) {
          T cp[3] = {w[0] - T(0.5) * dist * n[0], w[1] - T(0.5) * dist * n[1], w[2] - T(0.5) * dist * n[2]};
          dex_store_contact(stage + nslot * DEX_CON_STRIDE, dist, cp, n, p);
        } else stage[nslot * DEX_CON_STRIDE + 7] = T(-1);
      }
      nslot++;
    }
  }
  dex_syncwarp();
  // ---- append the env's work items to the global list (one atomic per env)
  int base = 0;
  if (lane == 0 && nwork > 0) base = dex_atomic_add_ret(st.nwork, nwork);
  base = dex_shfl(base, 0);
  bool ovf = false;
  for (int i = lane; i < nwork; i += 32) {
    const int wv = work[i];
    if (base + i < st.work_cap) { st.work[2 * (size_t)(base + i)] = env; st.work[2 * (size_t)(base + i) + 1] = wv; }
    else { ovf = true; stage[(wv >> 24) * DEX_CON_STRIDE + 7] = T(-1); }
  }
  if (dex_ballot(ovf)) status |= DEX_STATUS_CAND_OVERFLOW;
  if (lane == 0 && status) st.status[env] |= status;
  if (lane == 0) st.nslot[env] = nslot;
}

// K1c: staging slots -> contact list in slot order (warp per env).
// Returns the env's solver-cost class (0 = heaviest) for K2's processing order; the kernel wrapper files the env in
// the class list (one global atomic per class and CTA).
template <typename T>
DEX_DEV int dex_compact_env(const DevModel<T>& m, const BatchState<T>& st, int env) {
  if (dex_env_off(st, env)) return st.cls_mode ? -1 : DEX_NCLASS - 1;   // warp-uniform; with the two-kernel order a skipped env is simply not listed
  const int lane = dex_lane();
  const int nslot = st.nslot[env];
  const T* stage = st.stage + (size_t)env * DEX_STAGE_ENV;
  T* cout = st.contact + (size_t)env * DEX_MAXCON * DEX_CON_STRIDE;
  int ncon = 0;
  bool overflow = false, cross_bad = false;
  int ncross = 0;
  for (int s0 = 0; s0 < nslot; s0 += 32) {
    const int sl = s0 + lane;
    T rec[DEX_CON_STRIDE];
    bool valid = false;
    if (sl < nslot) {
      for (int k = 0; k < DEX_CON_STRIDE; ++k) rec[k] = stage[sl * DEX_CON_STRIDE + k];
      valid = rec[7] >= 0;
    }
    const unsigned vm = dex_ballot(valid);
    const int idx = ncon + dex_popc(vm & ((1u << lane) - 1u));
    bool cross = false;
    if (valid) {
      if (idx < DEX_MAXCON) { for (int k = 0; k < DEX_CON_STRIDE; ++k) cout[idx * DEX_CON_STRIDE + k] = rec[k]; }
      else overflow = true;
      if (st.cls_mode == 2) {   // sparse solver: the contact's dofs must lie on one root-to-leaf path of the chained elimination structure
        const int pr = (int)rec[7];
        const int b1 = dex_ldg(m.geom_body + dex_ldg(m.pair_geom1 + pr)), b2 = dex_ldg(m.geom_body + dex_ldg(m.pair_geom2 + pr));
        const unsigned mx = (unsigned)dex_ldg(m.body_dofmask + b1) | (unsigned)dex_ldg(m.body_dofmask + b2);
        cross_bad = cross_bad || ((mx & ~(unsigned)dex_ldg(m.body_augmask + b1)) && (mx & ~(unsigned)dex_ldg(m.body_augmask + b2)));
      } else if (st.cls_mode == 1) {   // does the contact join two branches of the arrowhead split (finger - finger)?
        const int pr = (int)rec[7];
        const unsigned mx = (unsigned)dex_ldg(m.body_dofmask + dex_ldg(m.geom_body + dex_ldg(m.pair_geom1 + pr))) ^
                            (unsigned)dex_ldg(m.body_dofmask + dex_ldg(m.geom_body + dex_ldg(m.pair_geom2 + pr)));
        int nbr = 0;
        for (int a = 0; a < 5; ++a) nbr += (mx & (unsigned)m.arrow_bmask[a]) ? 1 : 0;
        cross = nbr == 2 && !(mx & (unsigned)m.arrow_omask);   // the structured solver takes these as a low-rank correction
        cross_bad = cross_bad || (nbr > 1 && !cross);
      }
    }
    ncross += dex_popc(dex_ballot(cross));
    ncon += dex_popc(vm);
  }
  if (dex_ballot(overflow)) { if (lane == 0) st.status[env] |= DEX_STATUS_CON_OVERFLOW; }
  if (ncon > DEX_MAXCON) ncon = DEX_MAXCON;
  if (lane == 0) st.ncon[env] = ncon;
  // K2's processing order: the solver's cost grows with the contact count (mean Newton iterations per contact count:
  // 2.4, 3.5, 4.4, 5.2, 5.9, ...), so envs with contacts go first and the kernel does not end on a few long-running
  // warps (r01n: the SMs idled 14 % of K2's duration).  Keying on the previous step's iteration count instead was
  // slower: envs with equal contact counts also share code paths and cache lines.
  if (st.cls_mode == 2) {   // classes 0-3: dense form of the solver kernel, 4-7: sparse form; most contacts first within each
    if (dex_ballot(cross_bad)) return ncon >= 9 ? 0 : (ncon >= 4 ? 1 : (ncon >= 2 ? 2 : 3));
    return ncon >= 4 ? 4 : (ncon >= 2 ? 5 : (ncon >= 1 ? 6 : 7));
  }
  if (st.cls_mode == 1) {
    // Two solver kernels: classes 0-3 = dense warp-per-env solver (dex_dynamics.cuh), classes 4-7 = structured solver with
    // eight lanes per env (dex_arrow.cuh) for envs whose Hessian is block-arrowhead plus at most two contacts that join two
    // branches (a low-rank correction), with at most one contact per lane.  Within each kernel the envs with the most contacts go first; envs with equal contact
    // counts share a warp in the structured kernel (four envs per warp run as long as the slowest).
    const bool arrow = !dex_ballot(cross_bad) && ncross <= DEX_AR_MAXCROSS_K1 && ncon <= DEX_AR_MAXCON_K1;
    if (arrow) return ncross > 0 ? 4 : (ncon >= 3 ? 5 : (ncon >= 1 ? 6 : 7));   // envs with cross contacts share warps: their correction costs every env of the warp
    return ncon >= 12 ? 0 : (ncon >= 9 ? 1 : (ncon >= 4 ? 2 : 3));
  }
  const int key = ncon + 2;
  const int cls = key >= 9 ? 0 : (9 - key > DEX_NCLASS - 1 ? DEX_NCLASS - 1 : 9 - key);
  return cls;
}

#endif  // DEX_COLLIDE_CUH_
