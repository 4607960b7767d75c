// warp_emu.cpp — TEST INFRASTRUCTURE ONLY.
// A cooperative-fiber emulator of one CUDA warp so that the kernel bodies in
// dexterity_b200/csrc/*.cuh can be executed (slowly) on a host without a GPU, for debugging
// and for the CPU-only part of the test-suite.  It is NOT a fallback: the product library
// libdexb200.so is compiled by nvcc only, and nothing in the dexterity_b200 package loads
// this file.  32 ucontext fibers stand for the 32 lanes; every warp-synchronous primitive
// (shuffle, ballot, __syncwarp) is a fiber barrier.
#define DEX_WARP_EMU 1
#include <ucontext.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <thread>
#include <vector>

#include "../../dexterity_b200/csrc/dex_arrow.cuh"
#include "../../dexterity_b200/csrc/dex_collide.cuh"
#include "../../dexterity_b200/csrc/dex_dynamics.cuh"
#include "../../dexterity_b200/csrc/dex_sparse.cuh"
#include "../../dexterity_b200/csrc/dex_host_model.h"
#include "../../dexterity_b200/csrc/dex_task.cuh"
#include "../../dexterity_b200/csrc/dex_ik.cuh"
#include "../../dexterity_b200/csrc/dex_tree.cuh"

namespace dex_emu {
struct Warp {
  ucontext_t main_ctx, ctx[32];
  std::vector<char> stacks;
  int cur = 0;
  bool done[32];
  uint64_t xchg[32];
  bool pred[32];
  std::function<void()> fn;
};
static thread_local Warp* W = nullptr;
static const size_t kStack = 512 * 1024;

int lane() { return W->cur; }
void barrier() {
  int me = W->cur;
  swapcontext(&W->ctx[me], &W->main_ctx);
}
uint64_t exchange(uint64_t v, int src) {
  W->xchg[W->cur] = v;
  barrier();
  uint64_t r = W->xchg[src];
  barrier();
  return r;
}
unsigned ballot(bool p) {
  W->pred[W->cur] = p;
  barrier();
  unsigned r = 0;
  for (int l = 0; l < 32; ++l) if (!W->done[l] && W->pred[l]) r |= 1u << l;
  barrier();
  return r;
}
static void trampoline() {
  Warp* w = W;
  int me = w->cur;
  w->fn();
  w->done[me] = true;
  w->pred[me] = false;
  swapcontext(&w->ctx[me], &w->main_ctx);
}
void run_warp(const std::function<void()>& fn) {
  Warp* w = new Warp;
  W = w;
  w->fn = fn;
  w->stacks.resize(32 * kStack);
  for (int l = 0; l < 32; ++l) {
    w->done[l] = false; w->pred[l] = false; w->xchg[l] = 0;
    getcontext(&w->ctx[l]);
    w->ctx[l].uc_stack.ss_sp = w->stacks.data() + l * kStack;
    w->ctx[l].uc_stack.ss_size = kStack;
    w->ctx[l].uc_link = &w->main_ctx;
    makecontext(&w->ctx[l], trampoline, 0);
  }
  for (;;) {
    bool any = false;
    for (int l = 0; l < 32; ++l) {
      if (w->done[l]) continue;
      any = true;
      w->cur = l;
      swapcontext(&w->main_ctx, &w->ctx[l]);
    }
    if (!any) break;
  }
  W = nullptr;
  delete w;
}
}  // namespace dex_emu

// ------------------------------------------------------------------------------------------
template <typename T>
struct EmuModelT {
  DevModel<T> dev;
#define DEX_X_I(name) std::vector<int> v_##name;
#define DEX_X_R(name) std::vector<T> v_##name;
#define DEX_X(t, name) DEX_X_##t(name)
  DEX_DEVMODEL_ARRAYS(DEX_X)
#undef DEX_X
#undef DEX_X_I
#undef DEX_X_R
  void init(const HostModel& h) {
    dev.nbody = h.nbody; dev.njnt = h.njnt; dev.nq = h.nq; dev.nv = h.nv; dev.nu = h.nu; dev.ntendon = h.ntendon;
    dev.ngeom = h.ngeom; dev.npair = h.npair; dev.nsite = h.nsite; dev.nbodypair = h.nbodypair; dev.ntree = h.ntree;
    dev.k0_rounds = h.k0_rounds; dev.chol_split = h.chol_split; dev.maxdepth = h.maxdepth; dev.nconmax = h.nconmax; dev.iterations = h.iterations; dev.mpr_iterations = h.mpr_iterations; dev.nv_euler = h.nv_euler;
    dev.timestep = (T)h.timestep; dev.timestep_d = h.timestep;
    dev.arrow_ok = h.arrow_ok; dev.arrow_nmo = h.arrow_nmo; dev.arrow_omask = h.arrow_omask;
    dev.sp_ok = h.sp_ok; dev.sp[0] = h.sp[0]; dev.sp[1] = h.sp[1];
    std::memcpy(dev.arrow_bdof, h.arrow_bdof, sizeof(dev.arrow_bdof)); std::memcpy(dev.arrow_odof, h.arrow_odof, sizeof(dev.arrow_odof));
    std::memcpy(dev.arrow_bmask, h.arrow_bmask, sizeof(dev.arrow_bmask)); std::memcpy(dev.arrow_rowdof, h.arrow_rowdof, sizeof(dev.arrow_rowdof));
    for (int k = 0; k < 3; ++k) dev.gravity[k] = (T)h.gravity[k];
    dev.tolerance = (T)h.tolerance; dev.mpr_tolerance = (T)h.mpr_tolerance; dev.meaninertia = (T)h.meaninertia; dev.max_margin = (T)h.max_margin;
#define DEX_X_I(name) v_##name = h.name; dev.name = v_##name.data();
#define DEX_X_R(name) v_##name.assign(h.name.begin(), h.name.end()); dev.name = v_##name.data();
#define DEX_X(t, name) DEX_X_##t(name)
    DEX_DEVMODEL_ARRAYS(DEX_X)
#undef DEX_X
#undef DEX_X_I
#undef DEX_X_R
  }
};
struct EmuModel {
  HostModel host;
  EmuModelT<float> f;
  EmuModelT<double> d;
};

// collision stage of one env: broadphase (warp), narrowphase (one plain call per work item),
// compaction (warp).  The work list is private to the calling worker thread.
static std::atomic<long> g_arrow_envs{0}, g_dense_envs{0};
// DEX_EMU_SOLVER: 0 dense only, 1 (default, as in the library) contact-free envs structured (dex_arrow.cuh), 2 also the block-arrowhead envs with contacts,
// 3 the sparse form of the warp-per-env kernel (dex_sparse.cuh) for classes >= DEX_NCLASS_DENSE
template <typename T>
static int emu_solver_mode(const DevModel<T>& m) {
  const char* ms = getenv("DEX_EMU_SOLVER");
  int mode = getenv("DEX_EMU_DENSE") ? 0 : (ms ? atoi(ms) : 1);
  if ((mode == 1 || mode == 2) && !m.arrow_ok) mode = 0;
  if (mode == 3 && !m.sp_ok) mode = 0;
  return mode;
}
// returns the env's solver class (dex_compact_env): 4-7 = structured solver
template <typename T>
static int emu_collide(const DevModel<T>& m, const BatchState<T>& st, int env, T* smem) {
  std::vector<int> work(DEX_WORK_HDR + 2 * DEX_MAXSLOT, 0);
  BatchState<T> sl = st;
  sl.nwork = work.data(); sl.work = work.data() + DEX_WORK_HDR; sl.work_cap = DEX_MAXSLOT; sl.order = nullptr;
  sl.cls_mode = emu_solver_mode(m) == 3 ? 2 : (m.arrow_ok ? 1 : 0);
  dex_emu::run_warp([&]() { dex_collide_env<T>(m, sl, env, smem); });
  for (int i = 0; i < work[0]; ++i) dex_narrow_lane<T>(m, sl, work[DEX_WORK_HDR + 2 * i], work[DEX_WORK_HDR + 1 + 2 * i]);  // K1b is thread-per-item
  int cls = -1;
  dex_emu::run_warp([&]() { const int c = dex_compact_env<T>(m, sl, env); if (dex_lane() == 0) cls = c; });
  return cls;
}

// K2 for the envs e0 .. e1-1 with their solver classes: the structured solver for classes >= DEX_NCLASS_DENSE (one warp
// for the four envs, as on the GPU), the dense solver for the rest.  DEX_EMU_DENSE=1 forces the dense solver.
template <typename T>
static void emu_solve(const DevModel<T>& m, const BatchState<T>& st, const DynDebug<T>& dbg, int e0, int e1, const int* cls, T* smem,
                      int integrate) {
  int quad[4] = {-1, -1, -1, -1}, quadl[4] = {-1, -1, -1, -1};
  const int mode = dbg.M || dbg.efc ? 0 : emu_solver_mode(m);   // M / efc dumps: dense kernel only, as on the GPU
  bool any = false, anyl = false;
  for (int env = e0; env < e1; ++env) {
    const int c = cls[env - e0];
    if (mode == 3 && c >= DEX_NCLASS_DENSE) { dex_emu::run_warp([&]() { dex_dynamics_sp_env<T>(m, st, dbg, env, smem, integrate); }); g_arrow_envs++; }
    else if (mode >= 1 && mode <= 2 && c == DEX_NCLASS - 1) { quadl[env - e0] = env; anyl = true; g_arrow_envs++; }
    else if (mode == 2 && c >= DEX_NCLASS_DENSE) { quad[env - e0] = env; any = true; g_arrow_envs++; }
    else { dex_emu::run_warp([&]() { dex_dynamics_env<T>(m, st, dbg, env, smem, integrate); }); g_dense_envs++; }
  }
  if (any) dex_emu::run_warp([&]() { dex_arrow_warp<T, true>(m, st, dbg, quad[dex_lane() >> 3], smem, integrate); });
  if (anyl) dex_emu::run_warp([&]() { dex_arrow_warp<T, false>(m, st, dbg, quadl[dex_lane() >> 3], smem, integrate); });
}

// K0 for the envs e0 .. e0+3: the 8-lanes-per-env warp form when the model has a schedule, else one
// plain call per env (thread-per-env form)
template <typename T>
static void emu_tree(const DevModel<T>& m, const BatchState<T>& st, int e0, T* smem) {
  if (m.k0_rounds > 0 && !getenv("DEX_EMU_TREE_SCALAR")) dex_emu::run_warp([&]() { dex_tree_warp<T>(m, st, e0, smem); });
  else for (int env = e0; env < std::min(st.nenv, e0 + DEX_K0_ENVS); ++env) dex_tree_env<T>(m, st, env);
}

template <typename T>
static void emu_run(const DevModel<T>& m, BatchState<T> st, DynDebug<T> dbg, int nsub, int integrate, int do_collide,
                    int do_dynamics, int nthreads) {
  auto work = [&](int lo, int hi) {
    std::vector<T> smem(std::max<int>(std::max<int>(DEX_K1_SMEM_T, DEX_K2_SMEM_T), std::max<int>(DEX_K0_SMEM_T, std::max<int>(DEX_K2A_SMEM_T, DEX_K2S_SMEM_T))) + 64);
    BatchState<T> sb = st;
    sb.nenv = hi;   // the 4-env tree warps must not spill into another worker's range
    for (int e0 = lo; e0 < hi; e0 += DEX_K0_ENVS) {
      for (int s = 0; s < nsub; ++s) {
        emu_tree<T>(m, sb, e0, smem.data());
        int cls[DEX_K0_ENVS] = {0, 0, 0, 0};
        const int e1 = std::min(hi, e0 + DEX_K0_ENVS);
        for (int env = e0; env < e1; ++env) if (do_collide) cls[env - e0] = emu_collide<T>(m, sb, env, smem.data());
        if (do_dynamics) emu_solve<T>(m, sb, dbg, e0, e1, cls, smem.data(), integrate);
      }
    }
  };
  if (nthreads <= 1) { work(0, st.nenv); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) th.emplace_back(work, st.nenv * t / nthreads, st.nenv * (t + 1) / nthreads);
  for (auto& t : th) t.join();
}

extern "C" {
void* emu_model_create(const void* blob, uint64_t nbytes, char* err, int errlen) {
  EmuModel* m = new EmuModel;
  std::string e = dex_host::build(blob, nbytes, m->host);
  if (!e.empty()) { snprintf(err, errlen, "%s", e.c_str()); delete m; return nullptr; }
  m->f.init(m->host);
  m->d.init(m->host);
  return m;
}
void emu_model_destroy(void* m) { delete (EmuModel*)m; }
int emu_k2_smem_words() { return DEX_K2_SMEM_T; }
void emu_solver_counts(long* out, int reset) { out[0] = g_arrow_envs; out[1] = g_dense_envs; if (reset) { g_arrow_envs = 0; g_dense_envs = 0; } }
int emu_k0_rounds(void* m) { return ((EmuModel*)m)->host.k0_rounds; }

// Hill-climbing support function of K1b against an exhaustive scan, for every mesh geom and `ndir` random
// directions (warm-started from the previous direction's vertex, as inside one MPR run).
// out[0] = supports evaluated, out[1] = a different vertex was returned, out[2] = ... with a smaller support value
// (must be 0: on a convex hull every local maximum is global), out[3] = largest shortfall of the support value.
int emu_support_check(void* mp, int ndir, unsigned seed, double* out) {
  EmuModel* em = (EmuModel*)mp;
  const DevModel<double>& m = em->d.dev;
  unsigned s = seed ? seed : 1u;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return (double)(s & 0xffffff) / 0xffffff * 2.0 - 1.0; };
  out[0] = out[1] = out[2] = out[3] = 0;
  for (int g = 0; g < m.ngeom; ++g) {
    if (m.geom_type[g] != DEX_GEOM_MESH) continue;
    LaneGeom<double> G = {m.vert + 4 * m.geom_vertadr[g], m.geom_vertnum[g], DEX_GEOM_MESH, g, m.geom_adjbase[g], -1};
    double d[3] = {rnd(), rnd(), rnd()};
    for (int k = 0; k < ndir; ++k) {
      if (k % 8 == 0) { d[0] = rnd(); d[1] = rnd(); d[2] = rnd(); G.last = -1; }   // fresh direction, cold start
      else { d[0] += 0.3 * rnd(); d[1] += 0.3 * rnd(); d[2] += 0.3 * rnd(); }    // nearby direction, warm start
      double n = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      if (n < 1e-6) continue;
      const double l[3] = {d[0] / n, d[1] / n, d[2] / n};
      double got[3];
      dex_support_lane<double>(m, G, l[0], l[1], l[2], got);
      int bi = 0; double best = -1e300;
      for (int i = 0; i < G.n; ++i) {
        const double t = G.vert[4 * i] * l[0] + G.vert[4 * i + 1] * l[1] + G.vert[4 * i + 2] * l[2];
        if (t > best) { best = t; bi = i; }
      }
      const double tg = got[0] * l[0] + got[1] * l[1] + got[2] * l[2];
      out[0] += 1;
      if (got[0] != G.vert[4 * bi] || got[1] != G.vert[4 * bi + 1] || got[2] != G.vert[4 * bi + 2]) out[1] += 1;
      if (tg < best) { out[2] += 1; out[3] = std::max(out[3], best - tg); }
    }
  }
  return 0;
}

// dtype: 0 = float32, 1 = float64.  All arrays are host arrays of that dtype, row-major with the
// given strides.  dbg_* may be null.
int emu_run_steps(void* mp, int dtype, int nenv, void* qpos, int qpos_stride, void* qvel, int qvel_stride, void* ctrl,
                  int ctrl_stride, void* warm, int warm_stride, void* time, unsigned* status, void* contact, int* ncon,
                  int* niter, int nsub, int integrate, int do_collide, int do_dynamics, void* dbg_M, void* dbg_fs,
                  void* dbg_qacc, void* dbg_fc, void* dbg_efc, void* jc, void* mscr, void* xkin, void* tree, void* stage, int* nslot,
                  int nthreads) {
  EmuModel* m = (EmuModel*)mp;
  if (dtype == 0) {
    BatchState<float> st = {nenv, 0, (float*)qpos, qpos_stride, (float*)qvel, qvel_stride, (float*)ctrl, ctrl_stride,
                            (float*)warm, warm_stride, (float*)time, status, (float*)contact, ncon, niter, nullptr, (float*)jc, (float*)mscr,
                            (float*)xkin, (float*)stage, nslot, nullptr, nullptr, nullptr, 0, 0, (float*)tree};
    DynDebug<float> dbg = {(float*)dbg_M, (float*)dbg_fs, (float*)dbg_qacc, (float*)dbg_fc, (float*)dbg_efc};
    emu_run<float>(m->f.dev, st, dbg, nsub, integrate, do_collide, do_dynamics, nthreads);
  } else {
    BatchState<double> st = {nenv, 0, (double*)qpos, qpos_stride, (double*)qvel, qvel_stride, (double*)ctrl, ctrl_stride,
                             (double*)warm, warm_stride, (double*)time, status, (double*)contact, ncon, niter, nullptr, (double*)jc, (double*)mscr,
                             (double*)xkin, (double*)stage, nslot, nullptr, nullptr, nullptr, 0, 0, (double*)tree};
    DynDebug<double> dbg = {(double*)dbg_M, (double*)dbg_fs, (double*)dbg_qacc, (double*)dbg_fc, (double*)dbg_efc};
    emu_run<double>(m->d.dev, st, dbg, nsub, integrate, do_collide, do_dynamics, nthreads);
  }
  return 0;
}

// One batched control step (or a forced reset when force_reset != 0) through the task kernels.
// Pointers follow DexFieldId; `f` is an array of DEX_FIELD_COUNT host pointers, `stride` likewise.
int emu_env_step(void* mp, int dtype, int nenv, void** f, const int64_t* stride, const DexTaskParams* p, int nsub,
                 int force_reset, int nthreads) {
  EmuModel* m = (EmuModel*)mp;
  auto go = [&](auto tag, const auto& dev) {
    using T = decltype(tag);
    BatchState<T> st = {nenv, 0, (T*)f[DEX_FIELD_QPOS], (int)stride[DEX_FIELD_QPOS], (T*)f[DEX_FIELD_QVEL], (int)stride[DEX_FIELD_QVEL],
                        (T*)f[DEX_FIELD_CTRL], (int)stride[DEX_FIELD_CTRL], (T*)f[DEX_FIELD_QACC_WARMSTART],
                        (int)stride[DEX_FIELD_QACC_WARMSTART], (T*)f[DEX_FIELD_TIME], (unsigned*)f[DEX_FIELD_STATUS],
                        (T*)f[DEX_FIELD_CONTACT], (int*)f[DEX_FIELD_NCON], (int*)f[DEX_FIELD_NITER],
                        (const unsigned char*)f[DEX_FIELD_SKIP], (T*)f[DEX_FIELD_JC], (T*)f[DEX_FIELD_MSCR],
                        (T*)f[DEX_FIELD_XKIN], (T*)f[DEX_FIELD_STAGE], (int*)f[DEX_FIELD_NSLOT], nullptr, nullptr, nullptr, 0, 0,
                        (T*)f[DEX_FIELD_TREE]};
    TaskState<T> ts = {(const T*)f[DEX_FIELD_ACTION], (int)stride[DEX_FIELD_ACTION], (T*)f[DEX_FIELD_OBS], (T*)f[DEX_FIELD_REWARD],
                       (T*)f[DEX_FIELD_DISCOUNT], (T*)f[DEX_FIELD_GOAL], (int*)f[DEX_FIELD_TASK_I], (T*)f[DEX_FIELD_TASK_F],
                       (unsigned*)f[DEX_FIELD_RNG], (unsigned char*)f[DEX_FIELD_SKIP], (unsigned char*)f[DEX_FIELD_RESET_MASK],
                       (double*)f[DEX_FIELD_EPISODE_STATS]};
    st.time64 = (double*)f[DEX_FIELD_TIME64];
    DynDebug<T> dbg = {nullptr, nullptr, nullptr, nullptr, nullptr};
    auto work = [&](int lo, int hi) {
      std::vector<T> smem(std::max<int>(std::max<int>(std::max<int>(DEX_K1_SMEM_T, DEX_K2_SMEM_T), DEX_K3_SMEM_T), std::max<int>(DEX_K0_SMEM_T, std::max<int>(DEX_K2A_SMEM_T, DEX_K2S_SMEM_T))) + 64);
      BatchState<T> sb = st;
      sb.nenv = hi;
      for (int e0 = lo; e0 < hi; e0 += DEX_K0_ENVS) {
        const int e1 = std::min(hi, e0 + DEX_K0_ENVS);
        for (int env = e0; env < e1; ++env) dex_emu::run_warp([&]() { dex_task_pre_env<T>(dev, sb, ts, *p, env, force_reset); });
        if (p->task == DEX_TASK_REORIENT && p->place_max_attempts > 0) {   // PropPlacer rejection rounds (dex_place_env)
          BatchState<T> sp = sb;
          sp.skip_mode = 1;
          for (int r = 0; r < p->place_max_attempts; ++r) {
            bool any = false;
            for (int env = e0; env < e1; ++env) any = any || ts.skip[env] >= 2;
            if (!any) break;   // the GPU path launches the remaining rounds on no work
            emu_tree<T>(dev, sp, e0, smem.data());
            for (int env = e0; env < e1; ++env) {
              emu_collide<T>(dev, sp, env, smem.data());
              dex_emu::run_warp([&]() { dex_place_env<T>(dev, sp, ts, *p, env); });
            }
          }
        }
        if (!force_reset) for (int s = 0; s < nsub; ++s) {
          emu_tree<T>(dev, sb, e0, smem.data());
          int cls[DEX_K0_ENVS] = {0, 0, 0, 0};
          for (int env = e0; env < e1; ++env) cls[env - e0] = sb.skip && sb.skip[env] ? -1 : emu_collide<T>(dev, sb, env, smem.data());
          for (int env = e0; env < e1; ++env) if (cls[env - e0] < 0) cls[env - e0] = 0;   // skipped envs: the dense body returns at once
          emu_solve<T>(dev, sb, dbg, e0, e1, cls, smem.data(), 1);
        }
        for (int env = e0; env < e1; ++env) dex_emu::run_warp([&]() { dex_task_post_env<T>(dev, sb, ts, *p, env, smem.data()); });
      }
    };
    if (nthreads <= 1) { work(0, nenv); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work, nenv * t / nthreads, nenv * (t + 1) / nthreads);
    for (auto& t : th) t.join();
  };
  if (dtype == 0) go(float(0), m->f.dev); else go(double(0), m->d.dev);
  return 0;
}

int emu_ik_solve(void* mp, int dtype, const void* targets, int n, const DexIkParams* p, void* qout, int qstride,
                 unsigned char* ok, int* attempts, int nthreads) {
  EmuModel* m = (EmuModel*)mp;
  auto go = [&](auto tag, const auto& dev) {
    using T = decltype(tag);
    auto work = [&](int lo, int hi) {
      std::vector<T> smem(DEX_K6_SMEM_T + 64);
      for (int i = lo; i < hi; ++i)
        dex_emu::run_warp([&]() { dex_ik_env<T>(dev, *p, (const T*)targets, (T*)qout, qstride, ok, attempts, i, smem.data()); });
    };
    if (nthreads <= 1) { work(0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work, n * t / nthreads, n * (t + 1) / nthreads);
    for (auto& t : th) t.join();
  };
  if (dtype == 0) go(float(0), m->f.dev); else go(double(0), m->d.dev);
  return 0;
}
}
