// Microbenchmark: how fast can the SMs pull correlate-kernel tiles out of L2 / HBM?
//   rows tile  = 8 rows x 480 float2, contiguous 30720 B           (1-D bulk copy, UBLKCP)
//   cols tile  = 341 rows x 16 float2 at a row stride of 480 float2 (3-D tensor map, UTMALDG)
// against the same tiles fetched with plain LDG (128-bit for rows, 64-bit lanes-along-columns
// for columns). Consumers only sum what landed (LDS.128), so the figure is the load path's
// ceiling, not a kernel time. Build (no libcuda link: the encoder comes from the runtime):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tile_bw tools/microbench/tile_bw.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

constexpr int kRowsTileBytes = 8 * 480 * 8;
constexpr int kColsTileBytes = 341 * 16 * 8;

// MODE 0: 1-D bulk copies of contiguous tiles; MODE 1: 3-D tensor-map boxes (16 x 31 x 11)
template <int MODE, int STAGES, int THREADS>
__global__ void __launch_bounds__(THREADS) k_ring(const float2* __restrict__ src, const __grid_constant__ CUtensorMap map,
                                                  long long ntiles, int tiles_per_unit, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int TB = MODE == 0 ? kRowsTileBytes : kColsTileBytes;
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + (size_t)STAGES * TB);
  unsigned long long* empty = full + STAGES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], THREADS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](long long t, int s) {
    mbar_expect(&full[s], TB);
    if (MODE == 0) bulk_g2s(smem + (size_t)s * TB, reinterpret_cast<const unsigned char*>(src) + t * TB, TB, &full[s]);
    else { const int u = (int)(t / tiles_per_unit), c = (int)(t % tiles_per_unit); tma_3d(smem + (size_t)s * TB, &map, c * 16, 0, u * 11, &full[s]); }
  };
  float acc = 0.f;
  long long t = blockIdx.x;
  // prologue: fill the ring
  if (threadIdx.x == 0) {
    long long tt = t;
    for (int s = 0; s < STAGES && tt < ntiles; ++s, tt += gridDim.x) issue(tt, s);
  }
  int s = 0; unsigned ph = 0;
  for (; t < ntiles; t += gridDim.x) {
    mbar_wait(&full[s], ph);
    const float4* p = reinterpret_cast<const float4*>(smem + (size_t)s * TB);
    for (int i = threadIdx.x; i < TB / 16; i += THREADS) { const float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
    mbar_arrive(&empty[s]);
    if (threadIdx.x == 0) {
      const long long tn = t + (long long)STAGES * gridDim.x;
      if (tn < ntiles) { mbar_wait(&empty[s], ph); issue(tn, s); }
    }
    if (++s == STAGES) { s = 0; ph ^= 1; }
  }
  if (acc == 123.456f) out[0] = acc;
}

// plain loads: rows tiles with LDG.128 (UNROLL in flight per thread), columns tiles with 64-bit loads
template <int MODE, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS) k_ldg(const float2* __restrict__ src, long long ntiles, int tiles_per_unit, float* out) {
  float acc = 0.f;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    if (MODE == 0) {
      const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(src) + t * kRowsTileBytes);
      constexpr int n = kRowsTileBytes / 16;
      for (int i0 = threadIdx.x; i0 < n; i0 += THREADS * UNROLL) {
        float4 v[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) { const int i = i0 + k * THREADS; v[k] = i < n ? __ldg(&p[i]) : make_float4(0, 0, 0, 0); }
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
      }
    } else {
      const int u = (int)(t / tiles_per_unit), c = (int)(t % tiles_per_unit);
      const float2* p = src + (long long)u * 341 * 480 + c * 16 + (threadIdx.x & 15);
      for (int r0 = threadIdx.x >> 4; r0 < 341; r0 += (THREADS / 16) * UNROLL) {
        float2 v[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) { const int r = r0 + k * (THREADS / 16); v[k] = r < 341 ? __ldg(&p[(long long)r * 480]) : make_float2(0, 0); }
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) acc += v[k].x + v[k].y;
      }
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <class F> float time_ms(F&& f, int reps) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGetLastError());
  return ms / reps;
}

int main(int argc, char** argv) {
  const int units = argc > 1 ? atoi(argv[1]) : 32;          // 32 units x 1.31 MB = 42 MB: L2-resident when re-read
  const int reps = 20;
  const size_t nel = (size_t)units * 341 * 480;
  float2* src; float* out;
  CK(cudaMalloc(&src, nel * sizeof(float2))); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(src, 0, nel * sizeof(float2)));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  encode_fn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  CUtensorMap map;
  {
    // 64-bit elements: dims (480 columns, 31, 11*units), box (16, 31, 11) -> smem tile [11][31][16] = [341][16]
    cuuint64_t dims[3] = {480, 31, (cuuint64_t)11 * units};
    cuuint64_t strides[2] = {480 * 8, (cuuint64_t)480 * 8 * 31};
    cuuint32_t box[3] = {16, 31, 11};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
  }
  const double bytes = (double)nel * 8;
  const long long rows_tiles = (long long)nel * 8 / kRowsTileBytes;
  const int cols_tpu = 480 / 16;
  const long long cols_tiles = (long long)units * cols_tpu;
  printf("units=%d  bytes=%.1f MB  rows tiles=%lld  cols tiles=%lld  SMs=%d\n", units, bytes / 1e6, rows_tiles, cols_tiles, sms);
#define RUN_RING(MODE, ST, TH, CTAS)                                                                          \
  {                                                                                                           \
    auto kern = k_ring<MODE, ST, TH>;                                                                         \
    const size_t sm = (size_t)ST * (MODE == 0 ? kRowsTileBytes : kColsTileBytes) + 2 * ST * 8;                \
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));                      \
    const float ms = time_ms([&] { kern<<<sms * CTAS, TH, sm>>>(src, map, MODE == 0 ? rows_tiles : cols_tiles, cols_tpu, out); }, reps); \
    printf("%-10s stages=%d threads=%d ctas/sm=%d smem=%6zu  %8.3f ms  %8.1f GB/s\n", MODE == 0 ? "rows bulk" : "cols tma", ST, TH, CTAS, sm, ms, bytes / ms / 1e6); \
  }
#define RUN_LDG(MODE, UN, TH, CTAS)                                                                           \
  {                                                                                                           \
    auto kern = k_ldg<MODE, UN, TH>;                                                                          \
    const float ms = time_ms([&] { kern<<<sms * CTAS, TH>>>(src, MODE == 0 ? rows_tiles : cols_tiles, cols_tpu, out); }, reps); \
    printf("%-10s unroll=%d threads=%d ctas/sm=%d              %8.3f ms  %8.1f GB/s\n", MODE == 0 ? "rows ldg" : "cols ldg", UN, TH, CTAS, ms, bytes / ms / 1e6); \
  }
  RUN_RING(0, 2, 128, 1) RUN_RING(0, 3, 128, 1) RUN_RING(0, 4, 128, 1) RUN_RING(0, 6, 128, 1)
  RUN_RING(0, 2, 128, 2) RUN_RING(0, 3, 128, 2) RUN_RING(0, 2, 128, 3) RUN_RING(0, 2, 256, 1) RUN_RING(0, 4, 256, 1)
  RUN_RING(1, 2, 128, 1) RUN_RING(1, 3, 128, 1) RUN_RING(1, 4, 128, 1) RUN_RING(1, 2, 128, 2) RUN_RING(1, 2, 128, 4) RUN_RING(1, 2, 256, 2)
  RUN_LDG(0, 4, 128, 4) RUN_LDG(0, 8, 128, 4) RUN_LDG(0, 4, 128, 8) RUN_LDG(0, 8, 256, 4) RUN_LDG(0, 4, 256, 8)
  RUN_LDG(1, 4, 128, 4) RUN_LDG(1, 8, 128, 4) RUN_LDG(1, 11, 128, 4) RUN_LDG(1, 8, 128, 8) RUN_LDG(1, 11, 256, 4)
  return 0;
}
