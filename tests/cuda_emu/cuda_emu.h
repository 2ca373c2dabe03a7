// TEST HARNESS ONLY — a minimal host-side stand-in for the CUDA execution model.
//
// The build container has no GPU, and every gpurun round trip costs minutes, so the
// kernel sources under gnss-dsp-tools_b200/csrc are also compiled with g++ against
// this shim (tests/cuda_emu/build_emu.sh -> tests/cuda_emu/libgnssacq_emu.so) to check
// their index logic against the oracle on the CPU. One CUDA thread = one pooled
// std::thread, __syncthreads() = std::barrier, one block at a time. The emulated
// library is loaded only by tests (tests/emu_util.py); the product loader
// (gnsstools/_native.py) never looks for it.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define GNSSACQ_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
struct double2 { double x, y; };
struct int2 { int x, y; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline double2 make_double2(double a, double b) { return double2{a, b}; }

struct EmuIdx { unsigned x, y, z; };
extern thread_local EmuIdx threadIdx, blockIdx;
extern EmuIdx blockDim, gridDim;
extern std::barrier<>* emu_block_barrier;
extern unsigned char* emu_dyn_smem;
extern thread_local int emu_lane, emu_warp;
struct EmuWarp { uint64_t slot[32]; std::barrier<>* bar; };
extern EmuWarp* emu_warps;

static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __syncthreads() { emu_block_barrier->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warps[emu_warp].bar->arrive_and_wait(); }

template <class T> static inline T emu_shfl(T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  EmuWarp& w = emu_warps[emu_warp];
  uint64_t bits = 0; memcpy(&bits, &v, sizeof(T));
  w.slot[emu_lane] = bits;
  w.bar->arrive_and_wait();
  uint64_t got = w.slot[src & 31];
  w.bar->arrive_and_wait();
  T out; memcpy(&out, &got, sizeof(T));
  return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_shfl(v, emu_lane ^ m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) { return emu_shfl(v, emu_lane + d < 32 ? emu_lane + d : emu_lane); }
template <class T> static inline T __shfl_sync(unsigned, T v, int s) { return emu_shfl(v, s); }
static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned v = pred ? (1u << emu_lane) : 0u;
  for (int o = 16; o > 0; o >>= 1) v |= emu_shfl(v, emu_lane ^ o);
  return v;
}
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __any_sync(unsigned, int pred) {
  int v = pred ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) v |= emu_shfl(v, emu_lane ^ o);
  return v;
}

template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline long long __double2ll_rd(double a) { return (long long)std::floor(a); }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned long long atomicMax(unsigned long long* a, unsigned long long v) {
  unsigned long long old = __atomic_load_n(a, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
static inline unsigned atomicMax(unsigned* a, unsigned v) {
  unsigned old = __atomic_load_n(a, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
static inline float atomicAdd(float* a, float v) {
  static std::mutex mu; std::lock_guard<std::mutex> g(mu);
  float old = *a; *a = old + v; return old;
}
static inline int atomicAdd(int* a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_RELAXED); }
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }

// ---------------------------------------------------------------- runtime shim
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3,
       cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaStreamNonBlocking = 1 };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
// device memory is uninitialised on the GPU: poison it (0xff = NaN as float) so that reads of
// cells no kernel wrote show up as failures instead of lucky zeros
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); if (*p) memset(*p, 0xff, n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t = 0) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t hgt, int, cudaStream_t = 0) {
  for (size_t i = 0; i < hgt; ++i) memcpy(static_cast<char*>(d) + i * dp, static_cast<const char*>(s) + i * sp, w);
  return 0;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = malloc(1); return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = malloc(1); return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return 0; }
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; int l2CacheSize; char name[64]; };
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  p->multiProcessorCount = 4; p->sharedMemPerBlockOptin = 227 * 1024; p->l2CacheSize = 126 << 20; strcpy(p->name, "emu"); return 0; }

void emu_run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

// LAUNCH(kernel, grid, block, smem_bytes, stream, args...)
#define GNSSACQ_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu_run_grid((grid), (block), (smem), [&]() { kern(__VA_ARGS__); })
#define GNSSACQ_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu_dyn_smem)
#define __shared__ static
