// dex_api.cu — C ABI (include/dexterity_b200.h) + __global__ wrappers around the kernel bodies.
// Compiled by nvcc for sm_100a only; there is no host execution path in this library.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/dexterity_b200.h"
#include "dex_arrow.cuh"
#include "dex_collide.cuh"
#include "dex_dynamics.cuh"
#include "dex_sparse.cuh"
#include "dex_host_model.h"
#include "dex_ik.cuh"
#include "dex_task.cuh"
#include "dex_tree.cuh"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define DEX_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) return fail(DEX_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// Every entry point runs on the model's device and leaves the calling thread's current device as it found it
// (PyTorch keeps its own notion of the current device per thread).
struct DeviceGuard {
  int prev = -1; cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev); else if (err == cudaSuccess) prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define DEX_ON_DEVICE(dev)      \
  DeviceGuard guard_(dev);      \
  if (guard_.err != cudaSuccess) return fail(DEX_E_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(guard_.err))

#ifndef DEX_WARPS_PER_BLOCK
#define DEX_WARPS_PER_BLOCK 4
#endif
constexpr int kWarpsPerBlock = DEX_WARPS_PER_BLOCK;
#ifndef DEX_K1_MINB
#define DEX_K1_MINB 5
#endif
#ifndef DEX_K2_MINB
#define DEX_K2_MINB 5
#endif
#ifndef DEX_K2_CTAS
#define DEX_K2_CTAS 19   // register bound of K2 (65536 / (19 * 32) -> 96 registers, 21 CTAs resident, 12 bytes of spills instead of 48 at 80
                         // registers).  Re-tuned at the end of round 2, env step with 16 / 18 / 19 / 20 / 21 / 22 / 24: 3.63 / 3.59 / 3.57 / 3.59 /
                         // 3.64 / 3.67 / 3.65 ms (round 1, older kernel: 18: 0.449, 20: 0.451, 22: 0.421, 24: 0.432 ms per launch)
#endif

static_assert(DEX_STATUS_BIT_BADQACC == DEX_STATUS_BADQACC && DEX_STATUS_BIT_CON_OVERFLOW == DEX_STATUS_CON_OVERFLOW &&
              DEX_STATUS_BIT_CAND_OVERFLOW == DEX_STATUS_CAND_OVERFLOW && DEX_STATUS_BIT_SOLVER_MAXITER == DEX_STATUS_SOLVER_MAXITER &&
              DEX_STATUS_BIT_CHOL_FAIL == DEX_STATUS_CHOL_FAIL && DEX_STATUS_BIT_PLACEMENT == DEX_STATUS_PLACEMENT,
              "public status bits (include/dexterity_b200.h) out of step with the kernels'");

constexpr int kTreeBlock = 32;  // one env per THREAD; 16384 envs = 512 warps spread over all SMs
template <typename T>
__global__ void __launch_bounds__(kTreeBlock) k_tree(DevModel<T> m, BatchState<T> st) {
  const int env = st.env0 + blockIdx.x * kTreeBlock + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x == 0) { for (int k = 0; k < DEX_WORK_HDR; ++k) st.nwork[k] = 0; }  // work items, fetch cursor, order counters
  if (env >= st.nenv) return;
  dex_tree_env<T>(m, st, env);
}

#ifndef DEX_NARROW_BLOCK
#define DEX_NARROW_BLOCK 64
#endif
#ifndef DEX_TREE_WARPS
#define DEX_TREE_WARPS 1
#endif
constexpr int kTreeWarps = DEX_TREE_WARPS;  // K0, 8-lanes-per-env form: 4 envs per warp
#ifndef DEX_TREE_MINB
#define DEX_TREE_MINB 12   // register bound only (shared memory keeps 15 CTAs per SM): 109 registers, no spills; tree 0.132 -> 0.128 ms (15: 0.129, 18: 0.132, 20: 0.134)
#endif
template <typename T>
__global__ void __launch_bounds__(kTreeWarps * 32, DEX_TREE_MINB) k_tree8(DevModel<T> m, BatchState<T> st) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  if (blockIdx.x == 0 && threadIdx.x == 0) { for (int k = 0; k < DEX_WORK_HDR; ++k) st.nwork[k] = 0; }
  const int env_base = st.env0 + (blockIdx.x * kTreeWarps + warp) * DEX_K0_ENVS;
  if (env_base >= st.nenv) return;
  dex_tree_warp<T>(m, st, env_base, reinterpret_cast<T*>(smem_raw) + (size_t)warp * DEX_K0_SMEM_T);
}

constexpr int kNarrowBlock = DEX_NARROW_BLOCK;
// Persistent lanes: every thread runs one pair's MPR state machine and, as soon as that pair is done,
// takes the next item off the global list (warp-aggregated atomic on the cursor), so the 32 lanes of a
// warp stay busy although pairs need between 1 and ~50 support evaluations.
#ifndef DEX_NARROW_MINB
#define DEX_NARROW_MINB 8
#endif
#ifdef DEX_NARROW_PROF   // tuning aid (variants only): cycles and active lanes per trip index of a warp's loop
__device__ unsigned long long g_trip_cycles[128], g_trip_warps[128], g_trip_lanes[128];
#endif
template <typename T>
__global__ void __launch_bounds__(kNarrowBlock, DEX_NARROW_MINB) k_narrow(DevModel<T> m, BatchState<T> st) {
  int n = st.nwork[0];
#ifdef DEX_NARROW_PROF
  int trip = 0;
#endif
  if (n > st.work_cap) n = st.work_cap;
  const int lane = threadIdx.x & 31;
  MprLane<T> L;
  L.phase = MPR_DONE;
  bool have = false;
  for (;;) {
    if (L.phase == MPR_DONE) {
      if (have) L.finish(m, st);
      const unsigned am = __activemask();
      const int leader = __ffs(am) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(st.nwork + 1, __popc(am));
      base = __shfl_sync(am, base, leader);
      const int i = base + __popc(am & ((1u << lane) - 1u));
      if (i >= n) break;
      L.begin(m, st, st.work[2 * (size_t)i], st.work[2 * (size_t)i + 1]);
      have = true;
    }
#ifdef DEX_NARROW_PROF
    const long long t0 = clock64();
    const unsigned am0 = __activemask();
#endif
    L.step(m);
#ifdef DEX_NARROW_PROF
    {
      const long long t1 = clock64();
      const int tt = trip < 127 ? trip : 127;
      if (lane == __ffs(am0) - 1) { atomicAdd(&g_trip_cycles[tt], (unsigned long long)(t1 - t0)); atomicAdd(&g_trip_warps[tt], 1ull); atomicAdd(&g_trip_lanes[tt], (unsigned long long)__popc(am0)); }
      ++trip;
    }
#endif
  }
}

constexpr int kCompactWarps = 8;
template <typename T>
__global__ void __launch_bounds__(kCompactWarps * 32) k_compact(DevModel<T> m, BatchState<T> st) {
  __shared__ int cnt[DEX_NCLASS], base[DEX_NCLASS];
  const int lane = threadIdx.x & 31;
  const int env = st.env0 + blockIdx.x * kCompactWarps + (threadIdx.x >> 5);
  if (threadIdx.x < DEX_NCLASS) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int cls = env < st.nenv ? dex_compact_env<T>(m, st, env) : -1;
  if (!st.order) return;   // uniform over the grid
  // file the env in its class list: ranks within the CTA from a shared-memory counter, one global atomic per class
  int rank = 0;
  if (lane == 0 && cls >= 0) rank = atomicAdd(&cnt[cls], 1);
  __syncthreads();
  if (threadIdx.x < DEX_NCLASS && cnt[threadIdx.x] > 0) base[threadIdx.x] = atomicAdd(st.nwork + 2 + threadIdx.x, cnt[threadIdx.x]);
  __syncthreads();
  if (lane == 0 && cls >= 0) st.order[(size_t)cls * st.order_stride + st.env0 + base[cls] + rank] = env;
}

// K1a: one warp (= one env) per CTA: envs differ a lot in surviving pairs, single-warp CTAs retire independently
constexpr int kCollideWarps = 1;
template <typename T>
__global__ void __launch_bounds__(kCollideWarps * 32, DEX_K1_MINB * kWarpsPerBlock / kCollideWarps) k_collide(DevModel<T> m, BatchState<T> st) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int env = st.env0 + blockIdx.x * kCollideWarps + warp;
  if (env >= st.nenv) return;
  T* smem = reinterpret_cast<T*>(smem_raw) + (size_t)warp * DEX_K1_SMEM_T;
  dex_collide_env<T>(m, st, env, smem);
}

#ifndef DEX_K2_WARPS
#define DEX_K2_WARPS 1   // single-warp CTAs: envs differ 3x in solver work, a CTA of 4 would hold its slots until the slowest is done
#endif
constexpr int kDynWarps = DEX_K2_WARPS;
#ifndef DEX_K2S_CTAS
#define DEX_K2S_CTAS 22
#endif
// SP = 0: dense Cholesky form; SP = 1: sparse L^T D L form (dex_sparse.cuh).  The kernel takes K1c's classes [c0, c1).
template <typename T, int SP>
__global__ void __launch_bounds__(kDynWarps * 32, (SP ? DEX_K2S_CTAS : DEX_K2_CTAS) / DEX_K2_WARPS) k_dynamics(DevModel<T> m, BatchState<T> st, DynDebug<T> dbg,
                                                                   int integrate, int c0, int c1) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int slot = st.env0 + blockIdx.x * kDynWarps + warp;
  if (slot >= st.nenv) return;
  int env = slot;
  if (st.order) {   // K1c's lists, heaviest class first
    int s = slot - st.env0, c = c0;
    int cnt[DEX_NCLASS];   // all class counters in one go (a walk of dependent loads costs an L2 round trip per class)
#pragma unroll
    for (int k = 0; k < DEX_NCLASS; ++k) cnt[k] = st.nwork[2 + k];
#pragma unroll
    for (int k = 0; k < DEX_NCLASS; ++k) if (c == k && k < c1 && s >= cnt[k]) { s -= cnt[k]; ++c; }
    if (c >= c1) return;
    env = st.order[(size_t)c * st.order_stride + st.env0 + s];
  }
  T* smem = reinterpret_cast<T*>(smem_raw) + (size_t)warp * (SP ? DEX_K2S_SMEM_T : DEX_K2_SMEM_T);
  if (SP) dex_dynamics_sp_env<T>(m, st, dbg, env, smem, integrate);
  else dex_dynamics_env<T>(m, st, dbg, env, smem, integrate);
}

// K2A: structured solver, eight lanes per env.  Warp w takes the envs 4 w .. 4 w + 3 of the concatenated class lists
// DEX_NCLASS_DENSE .. DEX_NCLASS - 1 (most contacts first).
#ifndef DEX_K2A_CTAS
#define DEX_K2A_CTAS 8    // up to 255 registers: the solver state lives in registers; with more resident warps the spills and the lost L1 cost more than the residency gains (6 / 8 / 10 / 12 / 16 / 20 CTAs per SM: 0.156 / 0.157 / 0.173 / 0.175 / 0.215 / 0.283 ms per launch)
#endif
#ifndef DEX_K2L_CTAS
#define DEX_K2L_CTAS 12
#endif
// CON = true: classes [c0, c1) are envs with contacts (full form); CON = false: contact-free envs (class DEX_NCLASS - 1)
template <typename T, bool CON>
__global__ void __launch_bounds__(32, CON ? DEX_K2A_CTAS : DEX_K2L_CTAS) k_dynamics_arrow(DevModel<T> m, BatchState<T> st, DynDebug<T> dbg, int integrate,
                                                                                           int c0, int c1) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  int s = blockIdx.x * DEX_AR_ENVS + (lane >> 3), c = c0;
  int cnt[DEX_NCLASS];   // the class counters in one go (a walk of dependent loads costs an L2 round trip per class)
#pragma unroll
  for (int k = 0; k < DEX_NCLASS; ++k) cnt[k] = st.nwork[2 + k];
#pragma unroll
  for (int k = 0; k < DEX_NCLASS; ++k) if (c == k && k < c1 && s >= cnt[k]) { s -= cnt[k]; ++c; }
  const int env = c < c1 ? st.order[(size_t)c * st.order_stride + st.env0 + s] : -1;
  if (__ballot_sync(0xffffffffu, env >= 0) == 0) return;
  dex_arrow_warp<T, CON>(m, st, dbg, env, reinterpret_cast<T*>(smem_raw), integrate);
}

template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_task_pre(DevModel<T> m, BatchState<T> st, TaskState<T> ts,
                                                                   DexTaskParams p, int force_reset) {
  const int env = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (env >= st.nenv) return;
  dex_task_pre_env<T>(m, st, ts, p, env, force_reset);
}

// one rejection round of the collision-checked prop placement (dex_place_env)
template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_place(DevModel<T> m, BatchState<T> st, TaskState<T> ts, DexTaskParams p) {
  const int env = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (env >= st.nenv) return;
  dex_place_env<T>(m, st, ts, p, env);
}

template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_task_post(DevModel<T> m, BatchState<T> st, TaskState<T> ts,
                                                                    DexTaskParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int env = blockIdx.x * kWarpsPerBlock + warp;
  if (env >= st.nenv) return;
  T* smem = reinterpret_cast<T*>(smem_raw) + (size_t)warp * DEX_K3_SMEM_T;
  dex_task_post_env<T>(m, st, ts, p, env, smem);
}

template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_ik(DevModel<T> m, DexIkParams p, const T* targets, int n, T* qout,
                                                             int qstride, unsigned char* ok, int* attempts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int idx = blockIdx.x * kWarpsPerBlock + warp;
  if (idx >= n) return;
  T* smem = reinterpret_cast<T*>(smem_raw) + (size_t)warp * DEX_K6_SMEM_T;
  dex_ik_env<T>(m, p, targets, qout, qstride, ok, attempts, idx, smem);
}

template <typename T>
struct DeviceArrays {
  DevModel<T> dev;
  std::vector<void*> allocs;
  int upload(const HostModel& h) {
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
#define DEX_X_I(name)                                                                         \
  {                                                                                           \
    void* p = nullptr;                                                                        \
    size_t nb = std::max<size_t>(h.name.size(), 1) * sizeof(int);                             \
    DEX_CUDA(cudaMalloc(&p, nb));                                                             \
    allocs.push_back(p);                                                                      \
    if (!h.name.empty()) DEX_CUDA(cudaMemcpy(p, h.name.data(), h.name.size() * sizeof(int), cudaMemcpyHostToDevice)); \
    dev.name = (const int*)p;                                                                 \
  }
#define DEX_X_R(name)                                                                         \
  {                                                                                           \
    std::vector<T> tmp(h.name.begin(), h.name.end());                                         \
    void* p = nullptr;                                                                        \
    size_t nb = std::max<size_t>(tmp.size(), 1) * sizeof(T);                                  \
    DEX_CUDA(cudaMalloc(&p, nb));                                                             \
    allocs.push_back(p);                                                                      \
    if (!tmp.empty()) DEX_CUDA(cudaMemcpy(p, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice)); \
    dev.name = (const T*)p;                                                                   \
  }
#define DEX_X(t, name) DEX_X_##t(name)
    DEX_DEVMODEL_ARRAYS(DEX_X)
#undef DEX_X
#undef DEX_X_I
#undef DEX_X_R
    return DEX_OK;
  }
  void release() {
    for (void* p : allocs) cudaFree(p);
    allocs.clear();
  }
};

}  // namespace

#define DEX_MAXCHUNK 4

struct DexModel {
  int device = 0;
  HostModel host;
  DeviceArrays<float> f32;
  DeviceArrays<double> f64;
  bool has_f32 = false, has_f64 = false;
};

struct DexBatch {
  DexModel* model = nullptr;
  int nenv = 0;
  int dtype = DEX_F32;
  void* ptr[DEX_FIELD_COUNT];
  int64_t stride[DEX_FIELD_COUNT];
  int64_t launches = 0;
  bool attrs_set = false;
  // env chunks run as independent pipelines on library-owned streams (forked from / joined to the caller's)
  bool tree_scalar = false;  // force the thread-per-env tree kernel (tests / tuning)
  int solver_mode = 1;       // 0 dense solver only, 1 (default) structured solver for contact-free envs, 2 also for block-arrowhead envs with contacts,
                             // 3 sparse form of the warp-per-env kernel (dex_sparse.cuh) for every env whose contacts keep the Hessian in the chained pattern
  bool solver_serial = false;
  // the two solver kernels work on disjoint envs: the structured one runs on a second stream, forked after K1c and joined before the next K0
  cudaStream_t s2 = nullptr;
  cudaEvent_t ev_k1c = nullptr, ev_k2a = nullptr;
  int nchunk = 1;
  cudaStream_t cs[DEX_MAXCHUNK] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[DEX_MAXCHUNK] = {nullptr, nullptr, nullptr, nullptr};
  DexTaskParams task;
  bool task_set = false;
  // optional per-kernel-class timing (cudaEvent pairs recorded on the launching stream)
  bool profiling = false;
  std::vector<cudaEvent_t> ev[10];   // class: 0 pre, 1 broadphase, 2 dense solve, 3 post, 4 tree, 5 narrowphase, 6 compaction, 7 structured solve, 8 placement; pairs (start, stop)
  double ms[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  int64_t count[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
};

namespace {

struct ProfScope {  // records an event pair around one launch when profiling is on
  DexBatch* b; int cls; cudaStream_t s;
  ProfScope(DexBatch* b_, int cls_, cudaStream_t s_) : b(b_), cls(cls_), s(s_) {
    if (!b->profiling) return;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    b->ev[cls].push_back(e0); b->ev[cls].push_back(e1);
    cudaEventRecord(e0, s);
  }
  ~ProfScope() { if (b->profiling) cudaEventRecord(b->ev[cls].back(), s); }
};

// K1b runs a grid-stride loop over a device-side item count: size the grid for ~8 items per env,
// capped at a few CTAs per SM
inline int narrow_grid(int nenv) {
#ifndef DEX_NARROW_LANES_PER_ENV
#define DEX_NARROW_LANES_PER_ENV 6
#endif
  long long want = ((long long)nenv * DEX_NARROW_LANES_PER_ENV + kNarrowBlock - 1) / kNarrowBlock;
  const long long cap = 148 * 32;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

template <typename T>
int make_state(DexBatch* b, BatchState<T>& st, DynDebug<T>& dbg, bool need_dyn, bool use_skip = false) {
  // the tree kernel (K0) always runs, so its inputs and scratch are required by every entry point
  static const int req_col[] = {DEX_FIELD_QPOS, DEX_FIELD_STATUS, DEX_FIELD_CONTACT, DEX_FIELD_NCON, DEX_FIELD_QVEL, DEX_FIELD_CTRL,
                                DEX_FIELD_MSCR, DEX_FIELD_XKIN, DEX_FIELD_TREE, DEX_FIELD_STAGE, DEX_FIELD_NSLOT, DEX_FIELD_WORK};
  static const int req_dyn[] = {DEX_FIELD_QACC_WARMSTART, DEX_FIELD_TIME, DEX_FIELD_NITER, DEX_FIELD_JC};
  for (int f : req_col) if (!b->ptr[f]) return fail(DEX_E_UNBOUND, "required tensor not bound: field " + std::to_string(f));
  if (need_dyn) for (int f : req_dyn) if (!b->ptr[f]) return fail(DEX_E_UNBOUND, "required tensor not bound: field " + std::to_string(f));
  st.nenv = b->nenv;
  st.env0 = 0;
  st.qpos = (T*)b->ptr[DEX_FIELD_QPOS]; st.qpos_stride = (int)b->stride[DEX_FIELD_QPOS];
  st.qvel = (T*)b->ptr[DEX_FIELD_QVEL]; st.qvel_stride = (int)b->stride[DEX_FIELD_QVEL];
  st.ctrl = (T*)b->ptr[DEX_FIELD_CTRL]; st.ctrl_stride = (int)b->stride[DEX_FIELD_CTRL];
  st.warm = (T*)b->ptr[DEX_FIELD_QACC_WARMSTART]; st.warm_stride = (int)b->stride[DEX_FIELD_QACC_WARMSTART];
  st.time = (T*)b->ptr[DEX_FIELD_TIME];
  st.status = (unsigned*)b->ptr[DEX_FIELD_STATUS];
  st.contact = (T*)b->ptr[DEX_FIELD_CONTACT];
  st.ncon = (int*)b->ptr[DEX_FIELD_NCON];
  st.niter = (int*)b->ptr[DEX_FIELD_NITER];
  st.skip = use_skip ? (const unsigned char*)b->ptr[DEX_FIELD_SKIP] : nullptr;
  st.skip_mode = 0;
  st.time64 = (double*)b->ptr[DEX_FIELD_TIME64];
  st.jc = (T*)b->ptr[DEX_FIELD_JC];
  st.mscr = (T*)b->ptr[DEX_FIELD_MSCR];
  st.xkin = (T*)b->ptr[DEX_FIELD_XKIN];
  st.tree = (T*)b->ptr[DEX_FIELD_TREE];
  st.stage = (T*)b->ptr[DEX_FIELD_STAGE];
  st.nslot = (int*)b->ptr[DEX_FIELD_NSLOT];
  st.nwork = (int*)b->ptr[DEX_FIELD_WORK];
  st.work = st.nwork + DEX_WORK_HDR;
  st.work_cap = (int)b->stride[DEX_FIELD_WORK];
  st.order = (int*)b->ptr[DEX_FIELD_ORDER];
  st.order_stride = (int)b->stride[DEX_FIELD_ORDER];
  if (st.order && st.order_stride < b->nenv) return fail(DEX_E_ARG, "DEX_FIELD_ORDER must be [8][row_stride >= num_envs]");
  if (st.work_cap < 1) return fail(DEX_E_ARG, "DEX_FIELD_WORK must be bound with row_stride = capacity in items");
  dbg.M = (T*)b->ptr[DEX_FIELD_DBG_M];
  dbg.qfrc_smooth = (T*)b->ptr[DEX_FIELD_DBG_QFRC_SMOOTH];
  dbg.qacc = (T*)b->ptr[DEX_FIELD_DBG_QACC];
  dbg.qfrc_constraint = (T*)b->ptr[DEX_FIELD_DBG_QFRC_CONSTRAINT];
  dbg.efc = (T*)b->ptr[DEX_FIELD_DBG_EFC];
  return DEX_OK;
}

template <typename T>
int set_attrs(DexBatch* b) {
  if (b->attrs_set) return DEX_OK;
  DEX_CUDA(cudaFuncSetAttribute(k_collide<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kCollideWarps * DEX_K1_SMEM_T * sizeof(T))));
  DEX_CUDA(cudaFuncSetAttribute((k_dynamics<T, 0>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kDynWarps * DEX_K2_SMEM_T * sizeof(T))));
  DEX_CUDA(cudaFuncSetAttribute((k_dynamics<T, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kDynWarps * DEX_K2S_SMEM_T * sizeof(T))));
  DEX_CUDA(cudaFuncSetAttribute((k_dynamics_arrow<T, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)DEX_K2A_SMEM_T * sizeof(T))));
  DEX_CUDA(cudaFuncSetAttribute((k_dynamics_arrow<T, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)DEX_K2L_SMEM_T * sizeof(T))));
  // K0 / K1b keep per-thread state in local memory: give them the whole L1 (no shared-memory carve-out)
  DEX_CUDA(cudaFuncSetAttribute(k_tree<T>, cudaFuncAttributePreferredSharedMemoryCarveout, 0));
  DEX_CUDA(cudaFuncSetAttribute(k_tree8<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kTreeWarps * DEX_K0_SMEM_T * sizeof(T))));
  DEX_CUDA(cudaFuncSetAttribute(k_narrow<T>, cudaFuncAttributePreferredSharedMemoryCarveout, 0));
  b->attrs_set = true;
  return DEX_OK;
}

// nsub x (K0 tree, K1a broadphase, K1b narrowphase, K1c compaction, K2 solve) for every env.  With
// nchunk > 1 the env range is cut into contiguous chunks, each an independent pipeline on its own
// stream: the kernels are latency-bound at 20 % occupancy, so different stages of different chunks fill
// each other's idle issue slots and kernel tails.  Every chunk owns a slice of the work list.
template <typename T>
int launch_physics(DexBatch* b, const DevModel<T>& m, BatchState<T> st, const DynDebug<T>& dbg, int nsub, bool collide,
                   bool dynamics, int integrate, cudaStream_t s) {
  int rc = set_attrs<T>(b);
  if (rc) return rc;
  const size_t smem1 = (size_t)kCollideWarps * DEX_K1_SMEM_T * sizeof(T);
  const size_t smem2 = (size_t)kDynWarps * DEX_K2_SMEM_T * sizeof(T);
  const size_t smem2a = (size_t)DEX_K2A_SMEM_T * sizeof(T), smem2l = (size_t)DEX_K2L_SMEM_T * sizeof(T);
  int nchunk = b->nchunk;
  if (b->nenv < 1024 * nchunk) nchunk = 1;
  if (st.work_cap / nchunk < 64) nchunk = 1;
  if (nchunk > 1) {
    if (!b->ev_fork) {
      DEX_CUDA(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
      for (int c = 0; c < DEX_MAXCHUNK; ++c) {
        DEX_CUDA(cudaStreamCreateWithFlags(&b->cs[c], cudaStreamNonBlocking));
        DEX_CUDA(cudaEventCreateWithFlags(&b->ev_join[c], cudaEventDisableTiming));
      }
    }
    DEX_CUDA(cudaEventRecord(b->ev_fork, s));
  }
  int* const work_base = st.nwork;
  const int cap_total = st.work_cap;
  const int cap_c = (cap_total - DEX_WORK_HDR * (nchunk - 1)) / nchunk;  // items per chunk slice (each slice has its own header)
  for (int c = 0; c < nchunk; ++c) {
    cudaStream_t cs = s;
    if (nchunk > 1) { cs = b->cs[c]; DEX_CUDA(cudaStreamWaitEvent(cs, b->ev_fork, 0)); }
    st.env0 = (int)((long long)b->nenv * c / nchunk);
    st.nenv = (int)((long long)b->nenv * (c + 1) / nchunk);
    st.nwork = work_base + (size_t)c * (DEX_WORK_HDR + 2 * (size_t)cap_c);
    st.work = st.nwork + DEX_WORK_HDR;
    st.work_cap = cap_c;
    const int n = st.nenv - st.env0;
    const int tgrid = (n + kTreeBlock - 1) / kTreeBlock;
    const int t8grid = (n + kTreeWarps * DEX_K0_ENVS - 1) / (kTreeWarps * DEX_K0_ENVS);
    const size_t smem0 = (size_t)kTreeWarps * DEX_K0_SMEM_T * sizeof(T);
    const int ngrid = narrow_grid(n);
    st.cls_mode = (b->solver_mode == 3 && m.sp_ok) ? 2 : (m.arrow_ok ? 1 : 0);
    for (int i = 0; i < nsub; ++i) {
      {
        ProfScope ps(b, 4, cs);
        if (m.k0_rounds > 0 && !b->tree_scalar) k_tree8<T><<<t8grid, kTreeWarps * 32, smem0, cs>>>(m, st);
        else k_tree<T><<<tgrid, kTreeBlock, 0, cs>>>(m, st);
        b->launches++;
      }
      if (collide) {
        { ProfScope ps(b, 1, cs); k_collide<T><<<(n + kCollideWarps - 1) / kCollideWarps, kCollideWarps * 32, smem1, cs>>>(m, st); b->launches++; }
        { ProfScope ps(b, 5, cs); k_narrow<T><<<ngrid, kNarrowBlock, 0, cs>>>(m, st); b->launches++; }
        { ProfScope ps(b, 6, cs); k_compact<T><<<(n + kCompactWarps - 1) / kCompactWarps, kCompactWarps * 32, 0, cs>>>(m, st); b->launches++; }
      }
      if (dynamics) {
        // Solver kernels by K1c class (dex_compact_env): class 7 = envs without contacts -> structured solver, contact-free form;
        // classes 4-6 = block-arrowhead Hessian with contacts -> structured solver, full form (solver_mode 2) or dense;
        // classes 0-3 -> dense warp-per-env solver.  The kernels work on disjoint envs; the dense one runs on a second stream.
        int mode = (st.order && collide && !dbg.M && !dbg.efc) ? b->solver_mode : 0;
        if ((mode == 1 || mode == 2) && !m.arrow_ok) mode = 0;
        if (mode == 3 && !m.sp_ok) mode = 0;
        const int ndense = mode == 0 ? DEX_NCLASS : (mode == 1 ? DEX_NCLASS - 1 : DEX_NCLASS_DENSE);
        const int agrid = (n + DEX_AR_ENVS - 1) / DEX_AR_ENVS;
        const bool fork = mode != 0 && nchunk == 1 && !b->solver_serial;
        const size_t smem2s = (size_t)kDynWarps * DEX_K2S_SMEM_T * sizeof(T);
        cudaStream_t sa = cs;
        if (fork) {
          if (!b->s2) {
            DEX_CUDA(cudaStreamCreateWithFlags(&b->s2, cudaStreamNonBlocking));
            DEX_CUDA(cudaEventCreateWithFlags(&b->ev_k1c, cudaEventDisableTiming));
            DEX_CUDA(cudaEventCreateWithFlags(&b->ev_k2a, cudaEventDisableTiming));
          }
          DEX_CUDA(cudaEventRecord(b->ev_k1c, cs));
          DEX_CUDA(cudaStreamWaitEvent(b->s2, b->ev_k1c, 0));
          sa = b->s2;
        }
        if (mode == 2) { ProfScope ps(b, 7, sa); k_dynamics_arrow<T, true><<<agrid, 32, smem2a, sa>>>(m, st, dbg, integrate, DEX_NCLASS_DENSE, DEX_NCLASS - 1); b->launches++; }
        if (mode == 1 || mode == 2) { ProfScope ps(b, 9, sa); k_dynamics_arrow<T, false><<<agrid, 32, smem2l, sa>>>(m, st, dbg, integrate, DEX_NCLASS - 1, DEX_NCLASS); b->launches++; }
        if (mode == 3) { ProfScope ps(b, 7, sa); k_dynamics<T, 1><<<(n + kDynWarps - 1) / kDynWarps, kDynWarps * 32, smem2s, sa>>>(m, st, dbg, integrate, DEX_NCLASS_DENSE, DEX_NCLASS); b->launches++; }
        if (fork) DEX_CUDA(cudaEventRecord(b->ev_k2a, b->s2));
        { ProfScope ps(b, 2, cs); k_dynamics<T, 0><<<(n + kDynWarps - 1) / kDynWarps, kDynWarps * 32, smem2, cs>>>(m, st, dbg, integrate, 0, ndense); b->launches++; }
        if (fork) DEX_CUDA(cudaStreamWaitEvent(cs, b->ev_k2a, 0));
      }
    }
    if (nchunk > 1) { DEX_CUDA(cudaEventRecord(b->ev_join[c], cs)); DEX_CUDA(cudaStreamWaitEvent(s, b->ev_join[c], 0)); }
  }
  DEX_CUDA(cudaGetLastError());
  return DEX_OK;
}

template <typename T>
int run(DexBatch* b, const DevModel<T>& m, int nsub, bool collide, bool dynamics, int integrate, cudaStream_t s) {
  BatchState<T> st;
  DynDebug<T> dbg;
  int rc = make_state<T>(b, st, dbg, dynamics);
  if (rc) return rc;
  DEX_ON_DEVICE(b->model->device);
  return launch_physics<T>(b, m, st, dbg, nsub, collide, dynamics, integrate, s);
}

template <typename T>
int run_env(DexBatch* b, const DevModel<T>& m, int nsub, int force_reset, cudaStream_t s, int phases = 7) {
  if (!b->task_set) return fail(DEX_E_ARG, "dex_task_configure has not been called");
  static const int req[] = {DEX_FIELD_OBS, DEX_FIELD_REWARD, DEX_FIELD_DISCOUNT, DEX_FIELD_GOAL, DEX_FIELD_TASK_I,
                            DEX_FIELD_TASK_F, DEX_FIELD_RNG, DEX_FIELD_ACTION, DEX_FIELD_SKIP, DEX_FIELD_TIME64};
  for (int f : req) if (!b->ptr[f]) return fail(DEX_E_UNBOUND, "required task tensor not bound: field " + std::to_string(f));
  BatchState<T> st;
  DynDebug<T> dbg;
  int rc = make_state<T>(b, st, dbg, true, true);
  if (rc) return rc;
  TaskState<T> ts;
  ts.action = (const T*)b->ptr[DEX_FIELD_ACTION]; ts.action_stride = (int)b->stride[DEX_FIELD_ACTION];
  ts.obs = (T*)b->ptr[DEX_FIELD_OBS]; ts.reward = (T*)b->ptr[DEX_FIELD_REWARD]; ts.discount = (T*)b->ptr[DEX_FIELD_DISCOUNT];
  ts.goal = (T*)b->ptr[DEX_FIELD_GOAL]; ts.task_i = (int*)b->ptr[DEX_FIELD_TASK_I]; ts.task_f = (T*)b->ptr[DEX_FIELD_TASK_F];
  ts.rng = (unsigned*)b->ptr[DEX_FIELD_RNG]; ts.skip = (unsigned char*)b->ptr[DEX_FIELD_SKIP];
  ts.reset_mask = (unsigned char*)b->ptr[DEX_FIELD_RESET_MASK];
  ts.stats = (double*)b->ptr[DEX_FIELD_EPISODE_STATS];
  DEX_ON_DEVICE(b->model->device);
  const size_t smem3 = (size_t)kWarpsPerBlock * DEX_K3_SMEM_T * sizeof(T);
  const int grid = (b->nenv + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int block = kWarpsPerBlock * 32;
  if (phases & 1) { ProfScope ps(b, 0, s); k_task_pre<T><<<grid, block, 0, s>>>(m, st, ts, b->task, force_reset); b->launches++; }
  if ((phases & 1) && b->task.task == DEX_TASK_REORIENT && b->task.place_max_attempts > 0) {
    // PropPlacer's rejection loop for the envs reset above: collision kernels on those envs only, then one accept /
    // redraw decision per env and round.  Envs that are placed drop out; the launches of later rounds find no work.
    BatchState<T> sp = st;
    sp.skip_mode = 1;
    for (int r = 0; r < b->task.place_max_attempts; ++r) {
      rc = launch_physics<T>(b, m, sp, dbg, 1, true, false, 0, s);
      if (rc) return rc;
      ProfScope ps(b, 8, s);
      k_place<T><<<grid, block, 0, s>>>(m, sp, ts, b->task);
      b->launches++;
    }
  }
  if ((phases & 2) && !force_reset) {
    rc = launch_physics<T>(b, m, st, dbg, nsub, true, true, 1, s);
    if (rc) return rc;
  }
  if (phases & 4) { ProfScope ps(b, 3, s); k_task_post<T><<<grid, block, smem3, s>>>(m, st, ts, b->task); b->launches++; }
  DEX_CUDA(cudaGetLastError());
  return DEX_OK;
}

int dispatch_env(DexBatch* b, int nsub, int force_reset, void* stream, int phases = 7) {
  if (!b) return fail(DEX_E_ARG, "null batch");
  if (nsub < 1) return fail(DEX_E_ARG, "nsub must be >= 1");
  if (phases < 1 || phases > 7) return fail(DEX_E_ARG, "phases must be a non-empty subset of PRE|PHYSICS|POST");
  cudaStream_t s = (cudaStream_t)stream;
  if (b->dtype == DEX_F32) return run_env<float>(b, b->model->f32.dev, nsub, force_reset, s, phases);
  return run_env<double>(b, b->model->f64.dev, nsub, force_reset, s, phases);
}

int dispatch(DexBatch* b, int nsub, bool collide, bool dynamics, int integrate, void* stream) {
  if (!b) return fail(DEX_E_ARG, "null batch");
  if (nsub < 1) return fail(DEX_E_ARG, "nsub must be >= 1");
  cudaStream_t s = (cudaStream_t)stream;
  if (b->dtype == DEX_F32) return run<float>(b, b->model->f32.dev, nsub, collide, dynamics, integrate, s);
  return run<double>(b, b->model->f64.dev, nsub, collide, dynamics, integrate, s);
}

}  // namespace

extern "C" {

const char* dex_last_error(void) { return g_err.c_str(); }

int dex_capacity(int which) {
  switch (which) {
    case DEX_CAP_MAXCON: return DEX_MAXCON;
    case DEX_CAP_CON_STRIDE: return DEX_CON_STRIDE;
    case DEX_CAP_MAXDOF: return DEX_MAXDOF;
    case DEX_CAP_MAXBODY: return DEX_MAXBODY;
  }
  return -1;
}

int dex_model_create(const void* blob, uint64_t nbytes, int device, DexModel** out) {
  if (!blob || !out) return fail(DEX_E_ARG, "null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(DEX_E_CUDA, "no CUDA device: this library has no CPU execution path");
  if (device < 0 || device >= ndev) return fail(DEX_E_ARG, "bad device index");
  DexModel* m = new (std::nothrow) DexModel;
  if (!m) return fail(DEX_E_ARG, "out of host memory");
  std::string e = dex_host::build(blob, nbytes, m->host);
  if (!e.empty()) { delete m; return fail(DEX_E_MODEL, e); }
  m->device = device;
  *out = m;
  return DEX_OK;
}

void dex_model_destroy(DexModel* m) {
  if (!m) return;
  DeviceGuard guard_(m->device);
  m->f32.release();
  m->f64.release();
  delete m;
}

int dex_batch_create(DexModel* model, int num_envs, int dtype, DexBatch** out) {
  if (!model || !out || num_envs < 1) return fail(DEX_E_ARG, "bad argument");
  if (dtype != DEX_F32 && dtype != DEX_F64) return fail(DEX_E_ARG, "dtype must be DEX_F32 or DEX_F64");
  DEX_ON_DEVICE(model->device);
  if (dtype == DEX_F32 && !model->has_f32) { int rc = model->f32.upload(model->host); if (rc) return rc; model->has_f32 = true; }
  if (dtype == DEX_F64 && !model->has_f64) { int rc = model->f64.upload(model->host); if (rc) return rc; model->has_f64 = true; }
  DexBatch* b = new (std::nothrow) DexBatch;
  if (!b) return fail(DEX_E_ARG, "out of host memory");
  b->model = model; b->nenv = num_envs; b->dtype = dtype;
  std::memset(b->ptr, 0, sizeof(b->ptr));
  std::memset(b->stride, 0, sizeof(b->stride));
  *out = b;
  return DEX_OK;
}

void dex_batch_destroy(DexBatch* b) {
  if (!b) return;
  DeviceGuard guard_(b->model->device);
  for (auto& v : b->ev) { for (cudaEvent_t e : v) cudaEventDestroy(e); v.clear(); }   // profiling events not yet collected
  if (b->s2) { cudaStreamDestroy(b->s2); cudaEventDestroy(b->ev_k1c); cudaEventDestroy(b->ev_k2a); }
  if (b->ev_fork) {
    cudaEventDestroy(b->ev_fork);
    for (int c = 0; c < DEX_MAXCHUNK; ++c) { cudaStreamDestroy(b->cs[c]); cudaEventDestroy(b->ev_join[c]); }
  }
  delete b;
}

int dex_set_tree_kernel(DexBatch* b, int scalar) {
  if (!b) return fail(DEX_E_ARG, "null batch");
  b->tree_scalar = scalar != 0;
  return DEX_OK;
}

int dex_set_solver(DexBatch* b, int mode) {
  if (!b || mode < 0 || mode > 3) return fail(DEX_E_ARG, "solver mode must be 0, 1, 2 or 3");
  b->solver_mode = mode;
  b->solver_serial = getenv("DEX_SOLVER_SERIAL") != nullptr;   // profiling aid: all solver kernels on the caller's stream
  return DEX_OK;
}

int dex_set_chunks(DexBatch* b, int nchunk) {
  if (!b || nchunk < 1 || nchunk > DEX_MAXCHUNK) return fail(DEX_E_ARG, "nchunk must be in 1..4");
  b->nchunk = nchunk;
  return DEX_OK;
}

int dex_bind(DexBatch* b, int field, void* dev_ptr, int64_t row_stride) {
  if (!b || field < 0 || field >= DEX_FIELD_COUNT) return fail(DEX_E_ARG, "bad field");
  b->ptr[field] = dev_ptr;
  b->stride[field] = row_stride;
  return DEX_OK;
}

#ifdef DEX_EXP   // timing experiments (tools/k2_breakdown.py): DEX_FREEZE=1 computes the step but keeps the state
int dex_mj_step(DexBatch* b, int nsub, void* stream) { return dispatch(b, nsub, true, true, getenv("DEX_FREEZE") ? 2 : 1, stream); }
#else
int dex_mj_step(DexBatch* b, int nsub, void* stream) { return dispatch(b, nsub, true, true, 1, stream); }
#endif
int dex_mj_forward(DexBatch* b, void* stream) { return dispatch(b, 1, true, true, 0, stream); }
int dex_mj_collide(DexBatch* b, void* stream) { return dispatch(b, 1, true, false, 0, stream); }
int64_t dex_launch_count(const DexBatch* b) { return b ? b->launches : 0; }

int dex_task_configure(DexBatch* b, const DexTaskParams* p) {
  if (!b || !p) return fail(DEX_E_ARG, "null argument");
  const HostModel& h = b->model->host;
  if (p->task != DEX_TASK_REORIENT && p->task != DEX_TASK_REACH) return fail(DEX_E_ARG, "unknown task");
  if (p->nsite < 0 || p->nsite > 5 || p->nhand < 0 || p->nhand > h.nv) return fail(DEX_E_ARG, "bad nsite/nhand");
  for (int s = 0; s < p->nsite; ++s) if (p->site_id[s] < 0 || p->site_id[s] >= h.nsite) return fail(DEX_E_ARG, "bad site id");
  if (p->task == DEX_TASK_REORIENT) {
    if (p->prop_body <= 0 || p->prop_body >= h.nbody || p->prop_geom < 0 || p->prop_geom >= h.ngeom ||
        p->ground_geom < 0 || p->ground_geom >= h.ngeom) return fail(DEX_E_ARG, "bad prop/ground ids");
    if (3 * p->nhand + 6 * p->nsite + 21 > DEX_OBS_STRIDE) return fail(DEX_E_ARG, "observation does not fit");
    if (p->place_max_attempts < 0 || p->place_max_attempts > 200) return fail(DEX_E_ARG, "place_max_attempts must be in 0..200");
  }
  b->task = *p;
  b->task_set = true;
  return DEX_OK;
}
int dex_ik_solve(DexModel* model, int dtype, const void* targets, int n, const DexIkParams* p, void* qout,
                 int64_t row_stride, unsigned char* ok, int* attempts, void* stream) {
  if (!model || !targets || !p || !qout || !ok || n < 1) return fail(DEX_E_ARG, "bad argument");
  const HostModel& h = model->host;
  if (p->nsite < 1 || p->nsite > DEX_IK_MAXSITE) return fail(DEX_E_ARG, "nsite must be in 1..5");
  for (int s = 0; s < p->nsite; ++s) if (p->site_id[s] < 0 || p->site_id[s] >= h.nsite) return fail(DEX_E_ARG, "bad site id");
  if (h.nq != h.nv || h.njnt != h.nv) return fail(DEX_E_MODEL, "IK model must consist of hinge joints only");
  for (int j = 0; j < h.njnt; ++j)
    if (h.jnt_type[j] != DEX_JNT_HINGE || h.jnt_dofadr[j] != j || !h.jnt_limited[j]) return fail(DEX_E_MODEL, "IK model must consist of limited hinge joints");
  DEX_ON_DEVICE(model->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if (dtype == DEX_F32) {
    if (!model->has_f32) { int rc = model->f32.upload(h); if (rc) return rc; model->has_f32 = true; }
    const size_t smem = (size_t)kWarpsPerBlock * DEX_K6_SMEM_T * sizeof(float);
    DEX_CUDA(cudaFuncSetAttribute(k_ik<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_ik<float><<<grid, kWarpsPerBlock * 32, smem, st>>>(model->f32.dev, *p, (const float*)targets, n, (float*)qout, (int)row_stride, ok, attempts);
  } else if (dtype == DEX_F64) {
    if (!model->has_f64) { int rc = model->f64.upload(h); if (rc) return rc; model->has_f64 = true; }
    const size_t smem = (size_t)kWarpsPerBlock * DEX_K6_SMEM_T * sizeof(double);
    DEX_CUDA(cudaFuncSetAttribute(k_ik<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_ik<double><<<grid, kWarpsPerBlock * 32, smem, st>>>(model->f64.dev, *p, (const double*)targets, n, (double*)qout, (int)row_stride, ok, attempts);
  } else return fail(DEX_E_ARG, "bad dtype");
  DEX_CUDA(cudaGetLastError());
  return DEX_OK;
}

int dex_set_profiling(DexBatch* b, int on) {
  if (!b) return fail(DEX_E_ARG, "null batch");
  b->profiling = on != 0;
  return DEX_OK;
}
// Synchronises the device, folds the recorded event pairs into per-class totals and returns
// ms[10] (total milliseconds) and count[10] (launches) for pre / broadphase / dense solve / post / tree / narrowphase / compaction /
// structured solve (full form) / prop placement / structured solve (contact-free form).
int dex_get_kernel_times(DexBatch* b, double* ms, int64_t* count, int reset) {
  if (!b || !ms || !count) return fail(DEX_E_ARG, "null argument");
  DEX_ON_DEVICE(b->model->device);
  DEX_CUDA(cudaDeviceSynchronize());
  for (int c = 0; c < 10; ++c) {
    for (size_t i = 0; i + 1 < b->ev[c].size(); i += 2) {
      float t = 0;
      if (cudaEventElapsedTime(&t, b->ev[c][i], b->ev[c][i + 1]) == cudaSuccess) { b->ms[c] += t; b->count[c]++; }
      cudaEventDestroy(b->ev[c][i]); cudaEventDestroy(b->ev[c][i + 1]);
    }
    b->ev[c].clear();
    ms[c] = b->ms[c]; count[c] = b->count[c];
    if (reset) { b->ms[c] = 0; b->count[c] = 0; }
  }
  return DEX_OK;
}
#ifdef DEX_NARROW_PROF
int dex_debug_trips(unsigned long long* out, int reset) {   // out[3][128]: cycles, warps, lanes per trip index
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_trip_cycles, sizeof(g_trip_cycles));
  cudaMemcpyFromSymbol(out + 128, g_trip_warps, sizeof(g_trip_warps));
  cudaMemcpyFromSymbol(out + 256, g_trip_lanes, sizeof(g_trip_lanes));
  if (reset) { unsigned long long z[128] = {0}; cudaMemcpyToSymbol(g_trip_cycles, z, sizeof(z)); cudaMemcpyToSymbol(g_trip_warps, z, sizeof(z)); cudaMemcpyToSymbol(g_trip_lanes, z, sizeof(z)); }
  return 0;
}
#endif
int dex_env_reset(DexBatch* b, void* stream) { return dispatch_env(b, 1, 1, stream); }
int dex_env_step(DexBatch* b, int nsub, void* stream) { return dispatch_env(b, nsub, 0, stream); }
int dex_env_phase(DexBatch* b, int phases, int nsub, int force_reset, void* stream) { return dispatch_env(b, nsub, force_reset, stream, phases); }

}  // extern "C"
