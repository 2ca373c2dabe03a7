// Pipelined (persistent) correlate kernels for the plan-specialised large path.
//
// The one-tile-per-CTA kernels in kernels_spec.cuh start every tile with a burst of global
// loads that the whole CTA then waits on (ncu: 24-43 % of stall samples are long-scoreboard).
// Here a CTA stays resident, walks a list of tiles, and the bulk-copy engine (TMA,
// cp.async.bulk + mbarrier transaction counts) brings the *next* tile into a second
// shared-memory buffer while the butterflies of the current one run, so the loads of tile
// k+1 overlap the arithmetic of tile k. Arithmetic, data layout and results are those of the
// kernels_spec.cuh kernels (bit-identical: same operations in the same order).
//
// MEASURED SLOWER on B200 and therefore NOT part of libgnssacq.so (gnssacq.cu does not include
// this file): with <= 2 CTAs per SM the barrier waits that replace the load waits are no longer
// covered by other CTAs, and 372 separate 128-byte bulk copies per columns tile do not land in
// time (profiles/README.md, r02). Kept for tools/microbench/corr_pipe.cu; the small-CTA kernels
// in kernels_small.cuh are what came out of the comparison.
#pragma once
#include "../../gnss-dsp-tools_b200/csrc/kernels_small.cuh"

namespace acq {

// ---------------------------------------------------------------- async-copy primitives
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// every thread of the CTA arrives once per phase, announcing the bytes its own copies will deliver
__device__ __forceinline__ void mbar_arrive_expect(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy; bytes and both addresses are multiples of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// order earlier generic-proxy accesses of shared memory before later async-proxy writes
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#else
// host emulation (tests/cuda_emu): copies are synchronous, the wait is a block barrier
__device__ __forceinline__ void mbar_init(unsigned long long*, int) {}
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void mbar_arrive_expect(unsigned long long*, unsigned) {}
__device__ __forceinline__ void mbar_wait(unsigned long long*, unsigned) { __syncthreads(); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void fence_async_smem() {}
#endif

// =========================================================================== cols kernel
// grid = persistent CTAs (<= 2 per SM); items = (unit, column tile), item it = blockIdx.x +
// k*gridDim.x; each item is B steps (one per non-coherent block). Requires N2 even (16-byte
// aligned 128-byte tile rows). Shared memory: two tiles [N1][16] float2, q[N1][16] float when
// MULTI, two mbarriers.
template <class S, bool MULTI> __host__ __device__ constexpr size_t cols_pipe_smem() {
  return (size_t)S::F * kTileW * (2 * sizeof(float2) + (MULTI ? sizeof(float) : 0)) + 16;
}

template <class S, bool MULTI>
__global__ void __launch_bounds__(kThreads, 2)
k_corr_cols_p(DevPlan pl, const float2* __restrict__ scratch, int R_, int B, int D, int d0, int u0, int nunits,
              int n_lags, float scale, int ntiles, Part* __restrict__ parts, float* __restrict__ q_dump) {
  GNSSACQ_DYN_SMEM(float2, smem);
  constexpr int N1 = S::F, WP = kTileW, NS = S::NS;
  const int N = pl.N, N2 = pl.N2;
  float2* buf0 = smem;
  float2* buf1 = smem + N1 * WP;
  float* qs = reinterpret_cast<float*>(smem + 2 * N1 * WP);
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(smem) + cols_pipe_smem<S, MULTI>() - 16);
  const int tc = threadIdx.x & (kTW - 1);
  const int nitems = nunits * ntiles;

  if (threadIdx.x == 0) { mbar_init(&mbar[0], kThreads); mbar_init(&mbar[1], kThreads); mbar_fence_init(); }
  __syncthreads();

  auto issue = [&](int it, int b, int slot) {
    const int ul = it / ntiles, tl = it - ul * ntiles;
    const int col0 = tl * kTileW;
    const unsigned rowbytes = (unsigned)imin(kTileW, N2 - col0) * (unsigned)sizeof(float2);
    const float2* src = scratch + ((long long)ul * B + b) * N + col0;
    float2* dst = slot ? buf1 : buf0;
    constexpr int mine_max = (N1 + kThreads - 1) / kThreads;
    const int mine = ((int)threadIdx.x < N1 - (mine_max - 1) * kThreads) ? mine_max : mine_max - 1;
    mbar_arrive_expect(&mbar[slot], (unsigned)mine * rowbytes);
    for (int p1 = threadIdx.x; p1 < N1; p1 += kThreads) bulk_g2s(dst + p1 * WP, src + (long long)p1 * N2, rowbytes, &mbar[slot]);
  };

  int it = blockIdx.x, b = 0, step = 0;
  if (it < nitems) issue(it, 0, 0);
  float best = -1.f, sum = 0.f;
  int bestlag = 0x7fffffff;
  while (it < nitems) {
    int nit = it, nb_ = b + 1;
    if (nb_ == B) { nb_ = 0; nit = it + gridDim.x; }
    // the other buffer was last read by the previous step: every thread is past those reads
    __syncthreads();
    if (nit < nitems) { fence_async_smem(); issue(nit, nb_, (step + 1) & 1); }
    const int ul = it / ntiles, tl = it - ul * ntiles;
    const int col0 = tl * kTileW;
    const int ncols = imin(kTileW, N2 - col0);
    const int u = u0 + ul;
    const int r = u % R_, dd = u / R_;
    float* qd = q_dump ? q_dump + ((long long)r * D + d0 + dd) * N : nullptr;
    const int lag0 = col0 + tc;
    const bool last = (b + 1 == B);
    float2* tile = (step & 1) ? buf1 : buf0;
    mbar_wait(&mbar[step & 1], (unsigned)(step >> 1) & 1u);
    inv_stages_smem<S, NS - 1, 1, WP, 1>(tile, ncols, pl.s1);
    cols_last_stage<S, MULTI>(tile, qs, pl, ncols, lag0, b, last, n_lags, scale, qd, best, bestlag, sum);
    if (last) {
      unsigned long long key = bestlag != 0x7fffffff ? pack_key(best * scale, bestlag) : 0ull;
      sum *= scale;
      block_reduce_part(key, sum);
      if (threadIdx.x == 0) {
        Part p; p.key = key; p.sum = sum; p.pad = 0.f;
        parts[((long long)r * D + d0 + dd) * ntiles + tl] = p;
      }
      best = -1.f; sum = 0.f; bestlag = 0x7fffffff;
    }
    ++step; it = nit; b = nb_;
  }
}


// =========================================================================== rows kernel
// Persistent CTAs (<= 2 per SM). Items = (row tile of 8 rows, block b, unit), unit fastest; a
// CTA owns a contiguous run of items, so consecutive items usually share the row tile (the
// four-step twiddles it needs stay in L1). Shared memory: raw capture-spectrum rows Xs, raw
// replica-spectrum rows Cs (both filled by the bulk-copy engine one item ahead) and the working
// tile; pitch PP (float2) = 2 mod 4, so that 8 rows at one element offset cover all 32 banks
// with 16-byte accesses (quarter-warps) and, paired with a neighbouring element, with 8-byte
// accesses (half-warps). Thread = (row c = tid & 7, butterfly tid >> 3) except in the last
// stage, where lanes walk the row so that the global stores are contiguous.
//   [wait Xs,Cs] -> [Y = C conj(X) -> first inverse stage (unit stride) -> tile] -> barrier ->
//   [issue next item's copies] -> middle stages -> [last stage -> four-step twiddle -> store]
template <class S> __host__ __device__ constexpr size_t rows_pipe_smem() {
  return (size_t)3 * kRowsTile8 * rows8_pitch<S>() * sizeof(float2) + 16;
}

template <class S, int THREADS>
__global__ void __launch_bounds__(THREADS, 2)
k_corr_rows_p(DevPlan pl, const float2* __restrict__ X, const float2* __restrict__ C,
              int R_, int B, int u0, int nunits, float2* __restrict__ scratch) {
  GNSSACQ_DYN_SMEM(float2, smem);
  constexpr int N2 = S::F, PP = rows8_pitch<S>(), NS = S::NS, RT = kRowsTile8;
  static_assert(N2 % 2 == 0, "bulk copies need 16-byte rows");
  static_assert(THREADS % 32 == 0 && THREADS >= 64, "whole warps");
  const int N = pl.N, N1 = pl.N1;
  float2* Xs = smem;
  float2* Cs = smem + RT * PP;
  float2* tile = smem + 2 * RT * PP;
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(smem) + rows_pipe_smem<S>() - 16);
  const int nrt = (N1 + RT - 1) / RT;
  const int nitems = nrt * B * nunits;
  const int ipc = (nitems + (int)gridDim.x - 1) / (int)gridDim.x;
  int it = blockIdx.x * ipc;
  const int it_end = imin(nitems, it + ipc);

  if (threadIdx.x == 0) { mbar_init(&mbar[0], THREADS); mbar_fence_init(); }
  __syncthreads();

  constexpr unsigned rowbytes = (unsigned)N2 * (unsigned)sizeof(float2);
  auto issue = [&](int item) {
    const int ul = item % nunits, tb_ = item / nunits;
    const int b = tb_ % B, tl = tb_ / B;
    const int row0 = tl * RT, nrows = imin(RT, N1 - row0);
    const int u = u0 + ul, r = u % R_, dd = u / R_;
    const int c = threadIdx.x;
    if (c < nrows) {
      mbar_arrive_expect(&mbar[0], 2 * rowbytes);
      bulk_g2s(Xs + c * PP, X + ((long long)dd * B + b) * N + (long long)(row0 + c) * N2, rowbytes, &mbar[0]);
      bulk_g2s(Cs + c * PP, C + (long long)r * N + (long long)(row0 + c) * N2, rowbytes, &mbar[0]);
    } else {
      mbar_arrive_expect(&mbar[0], 0);
    }
  };

  if (it < it_end) issue(it);
  unsigned phase = 0;
  for (; it < it_end; ++it) {
    const int ul = it % nunits, tb_ = it / nunits;
    const int b = tb_ % B, tl = tb_ / B;
    const int row0 = tl * RT, nrows = imin(RT, N1 - row0);
    mbar_wait(&mbar[0], phase);
    phase ^= 1u;
    // ---- Y = C conj(X), first inverse stage (unit stride, no stage twiddle), into the tile
    {
      constexpr int R = S::radix(NS - 1), nbf = N2 / R;
      const int c = threadIdx.x & 7, tb = threadIdx.x >> 3;
      constexpr int nb = THREADS / 8;
      if (c < nrows) {
#pragma unroll 2
        for (int bf = tb; bf < nbf; bf += nb) {
          const int e0 = c * PP + bf * R;
          float2 v[R];
          if constexpr (R % 2 == 0) {
            const float4* xs4 = reinterpret_cast<const float4*>(Xs + e0);
            const float4* cs4 = reinterpret_cast<const float4*>(Cs + e0);
#pragma unroll
            for (int q = 0; q < R / 2; ++q) {
              const float4 cc = cs4[q], xx = xs4[q];
              v[2 * q] = cmulc(make_float2(cc.x, cc.y), make_float2(xx.x, xx.y));
              v[2 * q + 1] = cmulc(make_float2(cc.z, cc.w), make_float2(xx.z, xx.w));
            }
          } else {
#pragma unroll
            for (int q = 0; q < R; ++q) v[q] = cmulc(Cs[e0 + q], Xs[e0 + q]);
          }
          inv_dft<R>(v);
          if constexpr (R % 2 == 0) {
            float4* t4 = reinterpret_cast<float4*>(tile + e0);
#pragma unroll
            for (int q = 0; q < R / 2; ++q) t4[q] = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
          } else {
#pragma unroll
            for (int q = 0; q < R; ++q) tile[e0 + q] = v[q];
          }
        }
      }
    }
    __syncthreads();                       // Xs / Cs consumed, tile complete
    if (it + 1 < it_end) { fence_async_smem(); issue(it + 1); }
    inv_stages_rows8<S, NS - 2, PP, THREADS>(tile, nrows, pl.s2);
    rows8_last_stage<S, PP, THREADS>(tile, nrows, pl, scratch + ((long long)ul * B + b) * N + (long long)row0 * N2, row0);
    __syncthreads();                       // tile is rewritten by the next item's first stage
  }
}


}  // namespace acq
