// Small-CTA variants of the plan-specialised correlate kernels.
//
// ncu on the 256-thread kernels of kernels_spec.cuh shows issue slots ~40 % busy with the rest
// lost to global-load, barrier and dependency waits at 2-3 resident CTAs per SM. Halving the CTA
// (128-160 threads; 8-row tiles for the rows kernel) puts 4-7 independent CTAs on an SM, each at
// its own point of the load / butterfly / store sequence, and lets the radix-31 butterfly of the
// columns kernel live in one thread's registers instead of being shared by a warp pair (which
// duplicates its loads and twiddle multiplies). Same arithmetic in the same order as the
// kernels_spec.cuh kernels: results are bit-identical (tools/microbench/corr_pipe.cu checks it).
#pragma once
#include "kernels_spec.cuh"

namespace acq {

// Rows tile: 8 rows, pitch PP (float2) = 2 mod 4, so that 8 rows at one element offset cover all
// 32 banks with 16-byte accesses (quarter-warps) and, paired with a neighbouring element, with
// 8-byte accesses (half-warps). Thread = (row c = tid & 7, butterfly tid >> 3) except in the
// last stage, where lanes walk the row so that the global stores are contiguous.
constexpr int kRowsTile8 = 8;
#ifdef GNSSACQ_ROWS_LOAD_UNROLL
constexpr int kRowsLoadUnroll = GNSSACQ_ROWS_LOAD_UNROLL;     // A/B builds of tools/microbench only
#else
constexpr int kRowsLoadUnroll = 4;                            // 16-byte loads in flight per thread and operand
#endif
template <class S> __host__ __device__ constexpr int rows8_pitch() { return S::F + ((2 - S::F % 4) + 4) % 4; }

template <class S, int J, int PP, int THREADS>
__device__ __forceinline__ void inv_stage_rows8(float2* tile, int nrows, const float2* __restrict__ twbase, int twoff) {
  constexpr int R = S::radix(J), m = S::stride(J), nbf = S::F / R;
  static_assert(!is_split_radix(R), "warp-pair radices belong to the columns transform");
  const int c = threadIdx.x & 7, tb = threadIdx.x >> 3;
  constexpr int nb = THREADS / 8;
  const float2* tws = twbase + twoff;
  constexpr bool kHoist = !S::kPfa && m > 1 && nb % m == 0 && R <= 8;
  float2 wh[kHoist ? R : 1];
  if constexpr (kHoist) {
    const float2* w = tws + (tb % m) * (R - 1);
#pragma unroll
    for (int q = 1; q < R; ++q) wh[q] = __ldg(&w[q - 1]);
  }
  if (c < nrows) {
#pragma unroll stage_unroll(R)
    for (int bf = tb; bf < nbf; bf += nb) {
      const int blk = bf / m, i = bf - blk * m;
      float2* p = tile + c * PP + blk * R * m + i;
      float2 v[R];
#pragma unroll
      for (int q = 0; q < R; ++q) v[q] = p[q * m];
      if constexpr (kHoist) {
#pragma unroll
        for (int q = 1; q < R; ++q) v[q] = cmulc(v[q], wh[q]);
      } else if constexpr (m > 1 && !S::kPfa) {
        const float2* w = tws + i * (R - 1);
#pragma unroll
        for (int q = 1; q < R; ++q) v[q] = cmulc(v[q], __ldg(&w[q - 1]));
      }
      inv_dft<R>(v);
#pragma unroll
      for (int q = 0; q < R; ++q) p[q * m] = v[q];
    }
  }
  __syncthreads();
}
template <class S, int J, int PP, int THREADS>
__device__ __forceinline__ void inv_stages_rows8(float2* tile, int nrows, const SubPlan& sp) {
  if constexpr (J >= 1) {
    inv_stage_rows8<S, J, PP, THREADS>(tile, nrows, sp.tw, sp.tws_off[J]);
    inv_stages_rows8<S, J - 1, PP, THREADS>(tile, nrows, sp);
  }
}

// Last inverse stage of the rows transform (stage 0, stride m0 >= 16) fused with the conjugate
// four-step twiddle and the store: lanes walk i (consecutive positions), so tile reads, twiddle
// reads and global stores are contiguous. Prime-factor schedules have no stage twiddle here.
template <class S, int PP, int THREADS, bool GT = false>
__device__ __forceinline__ void rows8_last_stage(const float2* tile, int nrows, const DevPlan& pl, float2* __restrict__ out, int row0) {
  constexpr int N2 = S::F, R0 = S::radix(0), m0 = S::stride(0);
  static_assert(!is_split_radix(R0), "warp-pair radices belong to the columns transform");
  const float2* twm = pl.twm_inv + (long long)row0 * N2;
  const float2* twt = pl.s2.tw + pl.s2.tws0_t_off;
  const int items = m0 * nrows;
  for (int id = threadIdx.x; id < items; id += THREADS) {
    const int c = id / m0, i = id - c * m0;
    const float2* p = tile + c * PP + i;
    const int g = c * N2 + i;
    float2 v[R0];
#pragma unroll
    for (int q = 0; q < R0; ++q) v[q] = p[q * m0];
    if constexpr (!S::kPfa) {
#pragma unroll
      for (int q = 1; q < R0; ++q) v[q] = cmulc(v[q], __ldg(&twt[(q - 1) * m0 + i]));
    }
    inv_dft<R0>(v);
#pragma unroll
    for (int q = 0; q < R0; ++q) out[g + q * m0] = GT ? v[q] : cmulc(v[q], __ldg(&twm[g + q * m0]));
  }
}

// =========================================================================== rows kernel, small CTAs
// One 8-row tile per CTA, THREADS threads, no staging buffers: 28 KB of shared memory and <= 72
// registers let 6-7 CTAs share an SM, each at its own point of the load / butterfly / store
// sequence, so one CTA's global-load and barrier waits are covered by the others. Same
// arithmetic as k_corr_rows_s. grid = (ceil(N1/8), B, units).
template <class S> __host__ __device__ constexpr size_t rows_t_smem() {
  return (size_t)kRowsTile8 * rows8_pitch<S>() * sizeof(float2);
}
template <class S, int THREADS, int MINCTAS, bool GT = false>
__global__ void __launch_bounds__(THREADS, MINCTAS)
k_corr_rows_t(DevPlan pl, const float2* __restrict__ X, const float2* __restrict__ C,
              int R_, int B, int u0, float2* __restrict__ scratch) {
  GNSSACQ_DYN_SMEM(float2, tile);
  constexpr int N2 = S::F, PP = rows8_pitch<S>(), NS = S::NS, RT = kRowsTile8;
  static_assert(N2 % 2 == 0, "16-byte row accesses");
  const int N = pl.N, N1 = pl.N1;
  const int row0 = blockIdx.x * RT;
  const int nrows = imin(RT, N1 - row0);
  const int b = blockIdx.y, ul = blockIdx.z, u = u0 + ul;
  const int r = u % R_, dd = u / R_;
  const float4* Cr = reinterpret_cast<const float4*>(C + (long long)r * N + (long long)row0 * N2);
  const float4* Xb = reinterpret_cast<const float4*>(X + ((long long)dd * B + b) * N + (long long)row0 * N2);
  // ---- coalesced 16-byte loads, multiplied by the replica spectrum on the way in
  {
    constexpr int H2 = N2 / 2;
    float4* t4 = reinterpret_cast<float4*>(tile);
    const int n4 = nrows * H2;
#pragma unroll (kRowsLoadUnroll)
    for (int idx = threadIdx.x; idx < n4; idx += THREADS) {
      const int c = idx / H2, e2 = idx - c * H2;
      const float4 cc = __ldg(&Cr[idx]), xx = __ldg(&Xb[idx]);
      const float2 y0 = cmulc(make_float2(cc.x, cc.y), make_float2(xx.x, xx.y));
      const float2 y1 = cmulc(make_float2(cc.z, cc.w), make_float2(xx.z, xx.w));
      t4[c * (PP / 2) + e2] = make_float4(y0.x, y0.y, y1.x, y1.y);
    }
  }
  __syncthreads();
  // ---- first inverse stage (unit stride, no stage twiddle): 16-byte accesses, quarter-warp = 8 rows
  {
    constexpr int R = S::radix(NS - 1), nbf = N2 / R;
    const int c = threadIdx.x & 7, tb = threadIdx.x >> 3;
    constexpr int nb = THREADS / 8;
    if (c < nrows) {
#pragma unroll (R >= 16 ? 1 : 2)
      for (int bf = tb; bf < nbf; bf += nb) {
        const int e0 = c * PP + bf * R;
        float2 v[R];
        if constexpr (R % 2 == 0) {
          float4* t4 = reinterpret_cast<float4*>(tile + e0);
#pragma unroll
          for (int q = 0; q < R / 2; ++q) { const float4 t = t4[q]; v[2 * q] = make_float2(t.x, t.y); v[2 * q + 1] = make_float2(t.z, t.w); }
          inv_dft<R>(v);
#pragma unroll
          for (int q = 0; q < R / 2; ++q) t4[q] = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
        } else {
#pragma unroll
          for (int q = 0; q < R; ++q) v[q] = tile[e0 + q];
          inv_dft<R>(v);
#pragma unroll
          for (int q = 0; q < R; ++q) tile[e0 + q] = v[q];
        }
      }
    }
  }
  __syncthreads();
  inv_stages_rows8<S, NS - 2, PP, THREADS>(tile, nrows, pl.s2);
  rows8_last_stage<S, PP, THREADS, GT>(tile, nrows, pl, scratch + ((long long)ul * B + b) * N + (long long)row0 * N2, row0);
}


}  // namespace acq
