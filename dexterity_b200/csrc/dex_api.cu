// dex_api.cu — C ABI (include/dexterity_b200.h) + __global__ wrappers around the kernel bodies.
// Compiled by nvcc for sm_100a only; there is no host execution path in this library.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/dexterity_b200.h"
#include "dex_collide.cuh"
#include "dex_dynamics.cuh"
#include "dex_host_model.h"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define DEX_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) return fail(DEX_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

constexpr int kWarpsPerBlock = 4;

template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_collide(DevModel<T> m, BatchState<T> st) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int env = blockIdx.x * kWarpsPerBlock + warp;
  if (env >= st.nenv) return;
  T* smem = reinterpret_cast<T*>(smem_raw) + (size_t)warp * DEX_K1_SMEM_T;
  dex_collide_env<T>(m, st, env, smem);
}

template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_dynamics(DevModel<T> m, BatchState<T> st, DynDebug<T> dbg,
                                                                   int integrate) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int env = blockIdx.x * kWarpsPerBlock + warp;
  if (env >= st.nenv) return;
  T* smem = reinterpret_cast<T*>(smem_raw) + (size_t)warp * DEX_K2_SMEM_T;
  dex_dynamics_env<T>(m, st, dbg, env, smem, integrate);
}

template <typename T>
struct DeviceArrays {
  DevModel<T> dev;
  std::vector<void*> allocs;
  int upload(const HostModel& h) {
    dev.nbody = h.nbody; dev.njnt = h.njnt; dev.nq = h.nq; dev.nv = h.nv; dev.nu = h.nu; dev.ntendon = h.ntendon;
    dev.ngeom = h.ngeom; dev.npair = h.npair; dev.nsite = h.nsite; dev.nbodypair = h.nbodypair; dev.ntree = h.ntree;
    dev.maxdepth = h.maxdepth; dev.nconmax = h.nconmax; dev.iterations = h.iterations; dev.mpr_iterations = h.mpr_iterations;
    dev.timestep = (T)h.timestep;
    for (int k = 0; k < 3; ++k) dev.gravity[k] = (T)h.gravity[k];
    dev.tolerance = (T)h.tolerance; dev.mpr_tolerance = (T)h.mpr_tolerance; dev.meaninertia = (T)h.meaninertia;
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
};

namespace {

template <typename T>
int make_state(DexBatch* b, BatchState<T>& st, DynDebug<T>& dbg, bool need_dyn) {
  static const int req_col[] = {DEX_FIELD_QPOS, DEX_FIELD_STATUS, DEX_FIELD_CONTACT, DEX_FIELD_NCON};
  static const int req_dyn[] = {DEX_FIELD_QVEL, DEX_FIELD_CTRL, DEX_FIELD_QACC_WARMSTART, DEX_FIELD_TIME, DEX_FIELD_NITER};
  for (int f : req_col) if (!b->ptr[f]) return fail(DEX_E_UNBOUND, "required tensor not bound: field " + std::to_string(f));
  if (need_dyn) for (int f : req_dyn) if (!b->ptr[f]) return fail(DEX_E_UNBOUND, "required tensor not bound: field " + std::to_string(f));
  st.nenv = b->nenv;
  st.qpos = (T*)b->ptr[DEX_FIELD_QPOS]; st.qpos_stride = (int)b->stride[DEX_FIELD_QPOS];
  st.qvel = (T*)b->ptr[DEX_FIELD_QVEL]; st.qvel_stride = (int)b->stride[DEX_FIELD_QVEL];
  st.ctrl = (T*)b->ptr[DEX_FIELD_CTRL]; st.ctrl_stride = (int)b->stride[DEX_FIELD_CTRL];
  st.warm = (T*)b->ptr[DEX_FIELD_QACC_WARMSTART]; st.warm_stride = (int)b->stride[DEX_FIELD_QACC_WARMSTART];
  st.time = (T*)b->ptr[DEX_FIELD_TIME];
  st.status = (unsigned*)b->ptr[DEX_FIELD_STATUS];
  st.contact = (T*)b->ptr[DEX_FIELD_CONTACT];
  st.ncon = (int*)b->ptr[DEX_FIELD_NCON];
  st.niter = (int*)b->ptr[DEX_FIELD_NITER];
  dbg.M = (T*)b->ptr[DEX_FIELD_DBG_M];
  dbg.qfrc_smooth = (T*)b->ptr[DEX_FIELD_DBG_QFRC_SMOOTH];
  dbg.qacc = (T*)b->ptr[DEX_FIELD_DBG_QACC];
  dbg.qfrc_constraint = (T*)b->ptr[DEX_FIELD_DBG_QFRC_CONSTRAINT];
  dbg.efc = (T*)b->ptr[DEX_FIELD_DBG_EFC];
  return DEX_OK;
}

template <typename T>
int run(DexBatch* b, const DevModel<T>& m, int nsub, bool collide, bool dynamics, int integrate, cudaStream_t s) {
  BatchState<T> st;
  DynDebug<T> dbg;
  int rc = make_state<T>(b, st, dbg, dynamics);
  if (rc) return rc;
  DEX_CUDA(cudaSetDevice(b->model->device));
  const size_t smem1 = (size_t)kWarpsPerBlock * DEX_K1_SMEM_T * sizeof(T);
  const size_t smem2 = (size_t)kWarpsPerBlock * DEX_K2_SMEM_T * sizeof(T);
  if (!b->attrs_set) {
    DEX_CUDA(cudaFuncSetAttribute(k_collide<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    DEX_CUDA(cudaFuncSetAttribute(k_dynamics<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    b->attrs_set = true;
  }
  const int grid = (b->nenv + kWarpsPerBlock - 1) / kWarpsPerBlock;
  for (int i = 0; i < nsub; ++i) {
    if (collide) { k_collide<T><<<grid, kWarpsPerBlock * 32, smem1, s>>>(m, st); b->launches++; }
    if (dynamics) { k_dynamics<T><<<grid, kWarpsPerBlock * 32, smem2, s>>>(m, st, dbg, integrate); b->launches++; }
  }
  DEX_CUDA(cudaGetLastError());
  return DEX_OK;
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
  DEX_CUDA(cudaSetDevice(device));
  *out = m;
  return DEX_OK;
}

void dex_model_destroy(DexModel* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  m->f32.release();
  m->f64.release();
  delete m;
}

int dex_batch_create(DexModel* model, int num_envs, int dtype, DexBatch** out) {
  if (!model || !out || num_envs < 1) return fail(DEX_E_ARG, "bad argument");
  if (dtype != DEX_F32 && dtype != DEX_F64) return fail(DEX_E_ARG, "dtype must be DEX_F32 or DEX_F64");
  DEX_CUDA(cudaSetDevice(model->device));
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

void dex_batch_destroy(DexBatch* b) { delete b; }

int dex_bind(DexBatch* b, int field, void* dev_ptr, int64_t row_stride) {
  if (!b || field < 0 || field >= DEX_FIELD_COUNT) return fail(DEX_E_ARG, "bad field");
  b->ptr[field] = dev_ptr;
  b->stride[field] = row_stride;
  return DEX_OK;
}

int dex_mj_step(DexBatch* b, int nsub, void* stream) { return dispatch(b, nsub, true, true, 1, stream); }
int dex_mj_forward(DexBatch* b, void* stream) { return dispatch(b, 1, true, true, 0, stream); }
int dex_mj_collide(DexBatch* b, void* stream) { return dispatch(b, 1, true, false, 0, stream); }
int64_t dex_launch_count(const DexBatch* b) { return b ? b->launches : 0; }

}  // extern "C"
