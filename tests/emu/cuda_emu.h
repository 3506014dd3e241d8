// Host SIMT emulator for the CUDA sources of dada2_b200/csrc -- TEST INFRASTRUCTURE ONLY.
//
// tests/emu/build_emu.py rewrites the kernel launches of the .cu files textually and compiles them with g++ against
// this header into tests/emu/_build/libdada2b_emu.so.  The library exports the same C-ABI as libdada2b.so, so the CPU
// test-suite can run the *actual kernel sources* (warp shuffles, ballots, shared memory, atomics, the round driver)
// against the oracle without a GPU.  It is a checker for kernel logic, not a product path: nothing under dada2_b200/
// knows about it, it is orders of magnitude slower than the reference's CPU code, and it is never timed or shipped.
//
// Execution model: a launch runs its blocks one after another on the calling OS thread.  Every CUDA thread of the
// running block is a fiber (own stack, hand-written context switch); fibers run until they reach a warp- or block-wide
// collective (__shfl*_sync, __ballot_sync, __syncwarp, __syncthreads) and are resumed once every live thread of the warp
// (block) has arrived.  Exited threads leave the collectives, as on hardware.  All lanes of a warp must reach the same
// kind of collective in the same order -- a mismatch aborts with a message (it would be undefined behaviour on the GPU).
// Reading a shuffle value from an exited lane yields a poison pattern; cudaMalloc'ed memory and dynamic shared memory
// are poison-filled, so code that relies on zero-initialised memory fails here even where a fresh GPU would hide it.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <utility>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __constant__ static
#define __align__(n) alignas(n)

struct dim3 {
  unsigned x, y, z;
  constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

typedef struct cuemu_stream_st *cudaStream_t;
typedef struct cuemu_event_st *cudaEvent_t;
enum cudaError_t { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1 };

extern "C" const char *cuemu_self_path();
namespace cuemu {
extern thread_local uint3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
extern thread_local int t_lane;
void launch_impl(dim3 grid, dim3 block, size_t smem, void (*tramp)(void *), void *closure);
unsigned char *dyn_smem();
// Publishes one 64-bit value per lane and waits for the warp; returns the 32 published values and which lanes were live.
const uint64_t *warp_publish(uint64_t v, uint32_t *valid, int kind);
const uint64_t *warp_publish_masked(unsigned mask, uint64_t v, uint32_t *valid, int kind);
void block_barrier();
void *dev_alloc(size_t bytes);
void dev_free(void *p);
int num_sms();
enum { K_SHFL_UP = 1, K_SHFL_DOWN, K_SHFL_XOR, K_SHFL_IDX, K_BALLOT, K_ANY, K_SYNCWARP };

void note_launch(const char *kernel);
template <class F> void launch(const char *kernel, F f, dim3 grid, dim3 block, size_t smem = 0, cudaStream_t = nullptr) {
  note_launch(kernel);
  launch_impl(grid, block, smem, [](void *c) { (*static_cast<F *>(c))(); }, &f);
}
template <class T> inline uint64_t to_raw(T v) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  uint64_t r = 0;
  std::memcpy(&r, &v, sizeof(T));
  return r;
}
template <class T> inline T from_slot(const uint64_t *s, uint32_t valid, int src) {
  uint64_t r = ((valid >> src) & 1u) ? s[src] : 0xDEADBEEFDEADBEEFull;      // exited lane: undefined on hardware
  T v;
  std::memcpy(&v, &r, sizeof(T));
  return v;
}
inline void need_full(unsigned) {}      // partial masks: the lanes of the mask meet among themselves (warp_publish_masked)
}  // namespace cuemu

#define threadIdx (cuemu::t_threadIdx)
#define blockIdx (cuemu::t_blockIdx)
#define blockDim (cuemu::t_blockDim)
#define gridDim (cuemu::t_gridDim)
static const int warpSize = 32;

// ---- warp / block collectives ----
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  cuemu::need_full(mask);
  uint32_t valid;
  const uint64_t *s = cuemu::warp_publish_masked(mask, cuemu::to_raw(v), &valid, cuemu::K_SHFL_UP);
  const int lane = cuemu::t_lane, l = lane % width;
  return (l - (int)delta < 0) ? v : cuemu::from_slot<T>(s, valid, lane - (int)delta);
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  cuemu::need_full(mask);
  uint32_t valid;
  const uint64_t *s = cuemu::warp_publish_masked(mask, cuemu::to_raw(v), &valid, cuemu::K_SHFL_DOWN);
  const int lane = cuemu::t_lane, l = lane % width;
  return (l + (int)delta >= width) ? v : cuemu::from_slot<T>(s, valid, lane + (int)delta);
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
  cuemu::need_full(mask);
  uint32_t valid;
  const uint64_t *s = cuemu::warp_publish_masked(mask, cuemu::to_raw(v), &valid, cuemu::K_SHFL_XOR);
  const int lane = cuemu::t_lane, src = lane ^ lanemask;
  return (src / width != lane / width) ? v : cuemu::from_slot<T>(s, valid, src);
}
template <class T> inline T __shfl_sync(unsigned mask, T v, int srclane, int width = 32) {
  cuemu::need_full(mask);
  uint32_t valid;
  const uint64_t *s = cuemu::warp_publish_masked(mask, cuemu::to_raw(v), &valid, cuemu::K_SHFL_IDX);
  const int lane = cuemu::t_lane, src = (lane / width) * width + (((srclane % width) + width) % width);
  return cuemu::from_slot<T>(s, valid, src);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
  cuemu::need_full(mask);
  uint32_t valid;
  const uint64_t *s = cuemu::warp_publish_masked(mask, pred ? 1u : 0u, &valid, cuemu::K_BALLOT);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) if (((valid >> i) & 1u) && s[i]) r |= 1u << i;
  return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
inline void __syncwarp(unsigned mask = 0xffffffffu) {
  cuemu::need_full(mask);
  uint32_t valid;
  cuemu::warp_publish_masked(mask, 0, &valid, cuemu::K_SYNCWARP);
}
inline void __syncthreads() { cuemu::block_barrier(); }
inline void __threadfence() {}
inline void __threadfence_block() {}

using std::isfinite;
using std::isinf;
using std::isnan;

// ---- integer / bit intrinsics ----
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i); return r; }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) { return (unsigned)((((uint64_t)hi << 32) | lo) >> (shift & 31)); }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift) { return (unsigned)(((((uint64_t)hi << 32) | lo) << (shift & 31)) >> 32); }
inline int __vimax3_s32(int a, int b, int c) { return std::max(std::max(a, b), c); }
inline int __viaddmax_s32(int a, int b, int c) { return std::max((int)((unsigned)a + (unsigned)b), c); }   // VIADDMNMX: wrapping add, signed max
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * (uint64_t)b) >> 32); }
// 16-bit SIMD (two signed halves per word): VIADD.16x2 / VIMNMX.S16x2 with per-half predicate outputs / PRMT
inline unsigned __vadd2(unsigned a, unsigned b) { return ((a + b) & 0xFFFFu) | (((a >> 16) + (b >> 16)) << 16); }
inline unsigned __vibmax_s16x2(unsigned a, unsigned b, bool *pred_hi, bool *pred_lo) {
  const short al = (short)(a & 0xFFFFu), ah = (short)(a >> 16), bl = (short)(b & 0xFFFFu), bh = (short)(b >> 16);
  *pred_lo = al >= bl; *pred_hi = ah >= bh;
  return ((unsigned)(unsigned short)(al >= bl ? al : bl)) | ((unsigned)(unsigned short)(ah >= bh ? ah : bh) << 16);
}
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned sel) {
  const uint64_t v = ((uint64_t)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) { const unsigned n = (sel >> (4 * i)) & 0xFu; unsigned b = (unsigned)((v >> (8 * (n & 7))) & 0xFFu); if (n & 8) b = (b & 0x80u) ? 0xFFu : 0u; r |= b << (8 * i); }
  return r;
}
inline int __vimin3_s32(int a, int b, int c) { return std::min(std::min(a, b), c); }
inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long x) { double r; std::memcpy(&r, &x, 8); return r; }
inline int __float_as_int(float f) { int r; std::memcpy(&r, &f, 4); return r; }
inline float __int_as_float(int x) { float r; std::memcpy(&r, &x, 4); return r; }
template <class T> inline T __ldg(const T *p) { return *p; }

// CUDA's overloaded min/max (mixed signedness promotes like the usual arithmetic conversions)
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline std::common_type_t<A, B> min(A a, B b) { using C = std::common_type_t<A, B>; return (C)b < (C)a ? (C)b : (C)a; }
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline std::common_type_t<A, B> max(A a, B b) { using C = std::common_type_t<A, B>; return (C)a < (C)b ? (C)b : (C)a; }

// ---- atomics (blocks run one after another on one OS thread; the builtins keep this valid if that ever changes) ----
template <class T> inline std::enable_if_t<std::is_integral<T>::value, T> atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned v) { return __atomic_fetch_add(p, (unsigned long long)v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long *p, int v) { return __atomic_fetch_add(p, (unsigned long long)v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned *p, int v) { return __atomic_fetch_add(p, (unsigned)v, __ATOMIC_RELAXED); }
inline double atomicAdd(double *p, double v) { double o = *p; *p = o + v; return o; }
inline float atomicAdd(float *p, float v) { float o = *p; *p = o + v; return o; }
template <class T> inline T atomicSub(T *p, T v) { return __atomic_fetch_sub(p, v, __ATOMIC_RELAXED); }
template <class T> inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <class T> inline T atomicAnd(T *p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }
template <class T> inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
template <class T> inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicCAS(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

// ---- runtime API ----
template <class T> inline cudaError_t cudaMalloc(T **p, size_t bytes) { *p = (T *)cuemu::dev_alloc(bytes); return (*p || !bytes) ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> inline cudaError_t cudaMallocHost(T **p, size_t bytes) { *p = (T *)cuemu::dev_alloc(bytes); return (*p || !bytes) ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFree(void *p) { cuemu::dev_free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void *p) { cuemu::dev_free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)std::malloc(8); return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { return cudaStreamCreateWithFlags(s, 0); }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cuda_emu error"; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = cuemu::num_sms(); return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
struct cuemu_event_st { std::chrono::steady_clock::time_point t; };
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new cuemu_event_st(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
