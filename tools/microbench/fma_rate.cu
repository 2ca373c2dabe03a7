// Microbenchmark: scalar FFMA vs packed fma.rn.f32x2 issue rate on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu && ./fma_rate
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__global__ void k_ffma(float* out, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float* out, float a, float b) {
  unsigned long long x[16];
  unsigned long long aa, bb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
  for (int i = 0; i < 16; ++i) { float v = threadIdx.x * 0.001f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v)); }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// complex multiply-accumulate mix typical of the prime butterflies: re += c*a (both comps)
__global__ void k_mixed(float* out, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
  int k = threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = fmaf(x[i], a, b); k = k * 3 + i; }   // 1 FFMA + 1 IMAD
  }
  float s = k;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K> double run(K kern, const char* name, double flop_per_thread_iter) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<148 * 8, 256>>>(out, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) kern<<<148 * 8, 256>>>(out, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  double flops = 148.0 * 8 * 256 * ITERS * flop_per_thread_iter;
  printf("%-10s %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, ms, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  return ms;
}

int main() {
  run(k_ffma, "ffma", 16 * 2.0);
  run(k_ffma2, "ffma2", 16 * 4.0);
  run(k_mixed, "ffma+imad", 16 * 2.0);
  return 0;
}
