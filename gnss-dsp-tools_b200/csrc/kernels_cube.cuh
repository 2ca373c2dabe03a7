// N = 4096 = 16*16*16 (GPS L1 C/A and Xona X1 at 4.096 Msps: acquire-gps-l1.py:19-20): the whole
// transform lives in one CTA, 256 threads x 16 points, three radix-16 register butterflies
// with two shared-memory exchanges between them. Shared memory holds element idx at
// s[idx + idx/16] (one pad per 16) so all three access patterns are bank-conflict-free.
// Forward: DIF, natural in -> base-16 digit-reversed out (position 256*q0 + 16*q1 + q2 holds
// frequency q0 + 16*q1 + 256*q2); the correlate kernel is the exact adjoint and keeps the
// non-coherent sum of its 16 lags per thread in registers.
#pragma once
#include "kernels.cuh"

namespace acq {

constexpr int kCubeN = 4096;
constexpr int kCubeSmem = (kCubeN + kCubeN / 16) * (int)sizeof(float2);
__device__ __forceinline__ int cube_pad(int idx) { return idx + (idx >> 4); }

// tw0[(q-1)*256 + t] = exp(-2 pi i q t / 4096), q = 1..15, t < 256   (q-major: lanes read contiguously)
// tw1[(q-1)*16 + c]  = exp(-2 pi i q c / 256),  q = 1..15, c < 16
struct CubeTw { const float2* tw0; const float2* tw1; };

template <int SRC>
__global__ void __launch_bounds__(256, 3)
k_fwd_cube(CubeTw tw, const float2* __restrict__ x, const float* __restrict__ rep,
           const double* __restrict__ freq, const float2* __restrict__ nco_tab,
           int stride, int B, float2* __restrict__ X) {
  GNSSACQ_DYN_SMEM(float2, s);
  const int t = threadIdx.x;
  const int T = blockIdx.x;
  long long base;
  double f = 0.0;
  if (SRC == 0) { const int d = T / B, b = T - d * B; base = (long long)b * stride; f = freq[d]; }
  else { base = (long long)T * kCubeN; }
  float2 v[16];
  // stage 0: stride 256, fused with the load (+ carrier wipe-off)
#pragma unroll
  for (int a = 0; a < 16; ++a) v[a] = load_input<SRC>(x, rep, nco_tab, f, base, 256 * a + t);
  Dft<16>::run(v);
#pragma unroll
  for (int q = 1; q < 16; ++q) v[q] = cmul(v[q], __ldg(&tw.tw0[(q - 1) * 256 + t]));
#pragma unroll
  for (int q = 0; q < 16; ++q) s[cube_pad(256 * q + t)] = v[q];
  __syncthreads();
  // stage 1: stride 16 inside each block of 256
  {
    const int q0 = t >> 4, c = t & 15;
#pragma unroll
    for (int b = 0; b < 16; ++b) v[b] = s[cube_pad(256 * q0 + 16 * b + c)];
    Dft<16>::run(v);
#pragma unroll
    for (int q = 1; q < 16; ++q) v[q] = cmul(v[q], __ldg(&tw.tw1[(q - 1) * 16 + c]));
#pragma unroll
    for (int q = 0; q < 16; ++q) s[cube_pad(256 * q0 + 16 * q + c)] = v[q];
  }
  __syncthreads();
  // stage 2: unit stride
#pragma unroll
  for (int c = 0; c < 16; ++c) v[c] = s[17 * t + c];
  Dft<16>::run(v);
#pragma unroll
  for (int c = 0; c < 16; ++c) s[17 * t + c] = v[c];
  __syncthreads();
  float2* out = X + (long long)T * kCubeN;
#pragma unroll
  for (int k = 0; k < 16; ++k) out[256 * k + t] = s[cube_pad(256 * k + t)];
}

// grid.x = R * Dc; CTA (r, dd) loops over the B non-coherent blocks (as k_corr_mid).
template <bool MULTI>
__global__ void __launch_bounds__(256, 3)
k_corr_cube(CubeTw tw, const float2* __restrict__ X, const float2* __restrict__ C,
            int R, int B, int D, int d0, int n_lags, float scale,
            Part* __restrict__ parts, float* __restrict__ q_dump) {
  GNSSACQ_DYN_SMEM(float2, s);
  const int t = threadIdx.x;
  const int r = blockIdx.x % R, dd = blockIdx.x / R;
  const float2* Cr = C + (long long)r * kCubeN;
  float qacc[16];
#pragma unroll
  for (int a = 0; a < 16; ++a) qacc[a] = 0.f;
  float2 v[16];
  for (int b = 0; b < B; ++b) {
    const float2* Xb = X + ((long long)dd * B + b) * kCubeN;
    if (b > 0) __syncthreads();                    // previous block's stage 0 reads are done
#pragma unroll
    for (int k = 0; k < 16; ++k)
      s[cube_pad(256 * k + t)] = cmulc(__ldg(&Cr[256 * k + t]), __ldg(&Xb[256 * k + t]));
    __syncthreads();
    // adjoint of stage 2
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = cswap(s[17 * t + c]);
    Dft<16>::run(v);
#pragma unroll
    for (int c = 0; c < 16; ++c) s[17 * t + c] = cswap(v[c]);
    __syncthreads();
    // adjoint of stage 1
    {
      const int q0 = t >> 4, c = t & 15;
      v[0] = cswap(s[cube_pad(256 * q0 + c)]);
#pragma unroll
      for (int q = 1; q < 16; ++q)
        v[q] = cswap(cmulc(s[cube_pad(256 * q0 + 16 * q + c)], __ldg(&tw.tw1[(q - 1) * 16 + c])));
      Dft<16>::run(v);
#pragma unroll
      for (int bb = 0; bb < 16; ++bb) s[cube_pad(256 * q0 + 16 * bb + c)] = cswap(v[bb]);
    }
    __syncthreads();
    // adjoint of stage 0, outputs r[256 a + t] stay in registers
    v[0] = cswap(s[cube_pad(t)]);
#pragma unroll
    for (int q = 1; q < 16; ++q)
      v[q] = cswap(cmulc(s[cube_pad(256 * q + t)], __ldg(&tw.tw0[(q - 1) * 256 + t])));
    Dft<16>::run(v);                               // |.| below ignores the re/im swap
#pragma unroll
    for (int a = 0; a < 16; ++a) {
      const float mag = sqrtf(v[a].x * v[a].x + v[a].y * v[a].y);
      qacc[a] = MULTI ? qacc[a] + mag : mag;
    }
  }
  // peak / sum over this thread's 16 lags (ascending, so strict '>' keeps the first maximum)
  float best = -1.f, sum = 0.f;
  int bestlag = 0x7fffffff;
#pragma unroll
  for (int a = 0; a < 16; ++a) {
    const int lag = 256 * a + t;
    sum += qacc[a];
    if (lag < n_lags && qacc[a] > best) { best = qacc[a]; bestlag = lag; }
    if (q_dump) q_dump[((long long)r * D + d0 + dd) * kCubeN + lag] = qacc[a] * scale;
  }
  unsigned long long key = bestlag != 0x7fffffff ? pack_key(best * scale, bestlag) : 0ull;
  sum *= scale;
  block_reduce_part(key, sum);
  if (threadIdx.x == 0) {
    Part p; p.key = key; p.sum = sum; p.pad = 0.f;
    parts[(long long)r * D + d0 + dd] = p;
  }
}

}  // namespace acq
