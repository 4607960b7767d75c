// dex_portable.cuh — warp-level primitives used by every kernel body.
//
// Under nvcc these map 1:1 onto CUDA intrinsics (full-mask warp shuffles, ballots,
// __syncwarp, shared-memory atomics, __ldg).  When the header is compiled by a host
// compiler with DEX_WARP_EMU defined (tests/emu only — TEST INFRASTRUCTURE, never part of
// the product library) the same names are provided by a cooperative-fiber warp emulator
// so kernel bodies can be debugged without a GPU.  The product library
// (libdexb200.so) is compiled by nvcc only and contains no host execution path.
#ifndef DEX_PORTABLE_CUH_
#define DEX_PORTABLE_CUH_

#include <stdint.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define DEX_DEV __device__ __forceinline__
#define DEX_DEV_NOINLINE __device__ __noinline__
#define DEX_HD __host__ __device__ __forceinline__
#define DEX_FULL 0xffffffffu

DEX_DEV int dex_lane() { return threadIdx.x & 31; }
template <typename T> DEX_DEV T dex_shfl(T v, int src) { return __shfl_sync(DEX_FULL, v, src); }
template <typename T> DEX_DEV T dex_shfl_xor(T v, int m) { return __shfl_xor_sync(DEX_FULL, v, m); }
DEX_DEV unsigned dex_ballot(bool p) { return __ballot_sync(DEX_FULL, p); }
// shuffles inside a group of G consecutive lanes (G a power of two); groups of one warp may diverge
template <int G> DEX_DEV unsigned dex_gmask() { return G >= 32 ? DEX_FULL : (((1u << (G & 31)) - 1u) << (threadIdx.x & 31 & ~(G - 1))); }
template <int G, typename T> DEX_DEV T dex_gshfl_xor(T v, int m) { return __shfl_xor_sync(dex_gmask<G>(), v, m); }
DEX_DEV void dex_syncwarp() { __syncwarp(); }
template <typename T> DEX_DEV void dex_atomic_add(T* p, T v) { atomicAdd(p, v); }
DEX_DEV int dex_atomic_add_ret(int* p, int v) { return atomicAdd(p, v); }
template <typename T> DEX_DEV T dex_ldg(const T* p) { return __ldg(p); }
DEX_DEV int dex_popc(unsigned x) { return __popc(x); }
DEX_DEV int dex_ffs(unsigned x) { return __ffs(x); }

#elif defined(DEX_WARP_EMU)
#include <cmath>
#define DEX_DEV inline
#define DEX_DEV_NOINLINE inline
#define DEX_HD inline
namespace dex_emu {
int lane();
uint64_t exchange(uint64_t v, int src);  // shuffle: every live lane contributes v, reads lane src
unsigned ballot(bool p);
void barrier();
}  // namespace dex_emu
inline int dex_lane() { return dex_emu::lane(); }
template <typename T> inline T dex_shfl(T v, int src) {
  uint64_t raw = 0;
  __builtin_memcpy(&raw, &v, sizeof(T));
  raw = dex_emu::exchange(raw, src & 31);
  T out;
  __builtin_memcpy(&out, &raw, sizeof(T));
  return out;
}
template <typename T> inline T dex_shfl_xor(T v, int m) { return dex_shfl(v, dex_emu::lane() ^ m); }
inline unsigned dex_ballot(bool p) { return dex_emu::ballot(p); }
template <int G, typename T> inline T dex_gshfl_xor(T v, int m) { return dex_shfl(v, dex_emu::lane() ^ m); }
inline void dex_syncwarp() { dex_emu::barrier(); }
template <typename T> inline void dex_atomic_add(T* p, T v) { *p += v; }  // fibers are serialised
inline int dex_atomic_add_ret(int* p, int v) { int o = *p; *p += v; return o; }
template <typename T> inline T dex_ldg(const T* p) { return *p; }
inline int dex_popc(unsigned x) { return __builtin_popcount(x); }
inline int dex_ffs(unsigned x) { return __builtin_ffs((int)x); }
#else
#error "dex_portable.cuh: compile with nvcc (product) or -DDEX_WARP_EMU (tests/emu only)"
#endif

// 16-byte-aligned (x, y, z, pad) vertex load through the read-only path: one LDG.128 in FP32
#if defined(__CUDACC__)
DEX_DEV void dex_ldg_xyz(const float* p, float& x, float& y, float& z) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  x = v.x; y = v.y; z = v.z;
}
DEX_DEV void dex_ldg_xyz(const double* p, double& x, double& y, double& z) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(p));
  x = a.x; y = a.y; z = __ldg(p + 2);
}
#else
template <typename T> inline void dex_ldg_xyz(const T* p, T& x, T& y, T& z) { x = p[0]; y = p[1]; z = p[2]; }
#endif

// hint: bring the line at p into L1 (no register result, no stall)
#if defined(__CUDACC__)
DEX_DEV void dex_prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#else
inline void dex_prefetch_l1(const void*) {}
#endif

// four consecutive ints from a 16-byte-aligned address through the read-only path: one LDG.128
#if defined(__CUDACC__)
DEX_DEV void dex_ldg_int4(const int* p, int* out) {
  const int4 v = __ldg(reinterpret_cast<const int4*>(p));
  out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
#else
inline void dex_ldg_int4(const int* p, int* out) { out[0] = p[0]; out[1] = p[1]; out[2] = p[2]; out[3] = p[3]; }
#endif

// four consecutive reals from a 16-byte-aligned address through the read-only path (model constants)
#if defined(__CUDACC__)
DEX_DEV void dex_ldg4(const float* p, float* out) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
DEX_DEV void dex_ldg4(const double* p, double* out) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p + 2));
  out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
}
#else
template <typename T> inline void dex_ldg4(const T* p, T* out) { out[0] = p[0]; out[1] = p[1]; out[2] = p[2]; out[3] = p[3]; }
#endif

// four consecutive reals from a 16-byte (FP32) / 32-byte (FP64) aligned address: one LDG.128 in FP32
#if defined(__CUDACC__)
DEX_DEV void dex_ld4(const float* p, float* out) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
DEX_DEV void dex_ld4(const double* p, double* out) {
  const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
  out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
}
#else
template <typename T> inline void dex_ld4(const T* p, T* out) { out[0] = p[0]; out[1] = p[1]; out[2] = p[2]; out[3] = p[3]; }
#endif

// 16 consecutive reals to a 64-byte-aligned (FP32) / 128-byte-aligned (FP64) record: four STG.128 in FP32
#if defined(__CUDACC__)
DEX_DEV void dex_store16(float* p, const float* v) {
  float4* q = reinterpret_cast<float4*>(p);
  q[0] = make_float4(v[0], v[1], v[2], v[3]); q[1] = make_float4(v[4], v[5], v[6], v[7]);
  q[2] = make_float4(v[8], v[9], v[10], v[11]); q[3] = make_float4(v[12], v[13], v[14], v[15]);
}
DEX_DEV void dex_store16(double* p, const double* v) {
  double2* q = reinterpret_cast<double2*>(p);
#pragma unroll
  for (int k = 0; k < 8; ++k) q[k] = make_double2(v[2 * k], v[2 * k + 1]);
}
#else
template <typename T> inline void dex_store16(T* p, const T* v) { for (int k = 0; k < 16; ++k) p[k] = v[k]; }
#endif

// ---------------------------------------------------------------------------------------
// warp reductions (all 32 lanes must call)
// ---------------------------------------------------------------------------------------
template <typename T> DEX_DEV T dex_warp_sum(T v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += dex_shfl_xor(v, m);
  return v;
}
template <typename T> DEX_DEV T dex_warp_max(T v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) { T o = dex_shfl_xor(v, m); v = o > v ? o : v; }
  return v;
}
template <typename T> DEX_DEV T dex_warp_min(T v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) { T o = dex_shfl_xor(v, m); v = o < v ? o : v; }
  return v;
}
// argmax with lowest-index tie-break (matches a serial "first strict maximum" scan)
template <typename T> DEX_DEV void dex_warp_argmax(T& v, int& idx) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    T ov = dex_shfl_xor(v, m);
    int oi = dex_shfl_xor(idx, m);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
}
#if defined(__CUDACC__)
// FP32 specialisation: two REDUX instructions on an order-preserving integer key instead of ten
// dependent shuffles (the argmax latency dominated the MPR support function in the r01b profile).
template <> DEX_DEV void dex_warp_argmax<float>(float& v, int& idx) {
  const unsigned u = __float_as_uint(v);
  const unsigned key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  const unsigned mx = __reduce_max_sync(DEX_FULL, key);
  idx = __reduce_min_sync(DEX_FULL, key == mx ? idx : 0x7fffffff);
  const unsigned r = (mx & 0x80000000u) ? (mx & 0x7fffffffu) : ~mx;
  v = __uint_as_float(r);
}
#endif
// three independent warp sums with interleaved shuffles (hides the shuffle latency)
template <typename T> DEX_DEV void dex_warp_sum3(T& a, T& b, T& c) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    const T oa = dex_shfl_xor(a, m), ob = dex_shfl_xor(b, m), oc = dex_shfl_xor(c, m);
    a += oa; b += ob; c += oc;
  }
}

#endif  // DEX_PORTABLE_CUH_
