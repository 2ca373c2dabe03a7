// Acquisition kernels: carrier wipe-off + forward FFT of the capture blocks, and the fused
// correlate kernel (replica-spectrum multiply -> inverse FFT -> |.| -> non-coherent sum ->
// max / argmax / sum), in a resident ("mid", one CTA per transform) and a two-kernel
// ("large", through an L2-resident scratch) form. See DESIGN.md for the data layout.
//
// Reference behaviour being reproduced (pmonta/GNSS-DSP-tools):
//   w = nco.nco(-doppler/fs, 0, N)            gnsstools/nco.py:6-10, acquire-gps-l1.py:28
//   b = x[block*n : block*n+N] * w            acquire-gps-l1.py:30-31, acquire-gps-l5i.py:30-31
//   r = ifft(C * conj(fft(b)))                acquire-gps-l1.py:32
//   q += |r|                                  acquire-gps-l1.py:33
//   idx = argmax(q); metric = q[idx](/mean q) acquire-gps-l1.py:34-35, acquire-gps-l5i.py:34-36
#pragma once
#include "fft_core.cuh"

namespace acq {

constexpr int kThreads = 256;
constexpr int kNcoSize = 1024;

struct DevPlan {
  int N, N1, N2;
  SubPlan s1, s2;
  const float2* twm;     // twm[p1*N2 + n2]: forward four-step twiddle (columns = natural n2)
  const float2* twm_inv; // [p1*N2 + p2]: the same for the inverse, whose columns are tile positions (== twm unless s2.pfa)
  const int* n1_of_pos;  // time index n1 held at position p1 of the length-N1 tile (identity unless s1.pfa)
  const int* n2_of_pos;  // likewise for the length-N2 tile
  const int* pos2_of_n;  // inverse of n2_of_pos
  // Coprime split (fft_plan.h, HostPlan::gt): no four-step twiddles; sample n = m*N2 + b goes to tile
  // position fpos1[n mod N1] of column b, a stored row goes through fpos2[b], and the lag of inverse
  // output (n1, column q2) is (n1*N2 + col_lag[q2]) mod N. col_lag is valid for every plan
  // (n2_of_pos when gt == 0).
  int gt;
  const int* fpos1;
  const int* fpos2;
  const int* col_lag;
  // Embedded lengths (gnssacq.cu, any N the planner cannot factor runs inside a longer power-of-two
  // transform): capture samples at block offsets >= xlen read as zero and only lags < sum_lags enter
  // the sum behind the mean. Both equal N otherwise; honoured by the generic kernels, which are the
  // ones an embedded plan runs.
  int xlen, sum_lags;
};

// Units of one launch of the kernels_v3.cuh pair: replicas [r0, r0+Rc) x chunk-local Doppler bins
// [dd0, dd0+G) (x B blocks); scratch slot of (replica r, bin dd, block b) below.
struct ChunkV3 { int r0, Rc, dd0, G; };
__host__ __device__ __forceinline__ int v3_slot(const ChunkV3& ck, int B, int r, int dd, int b) {
  return ((r - ck.r0) * ck.G + (dd - ck.dd0)) * B + b;
}

// One launch of the fused persistent correlate kernel (kernels_fused.cuh) and its global counters.
struct FusedJob {
  int R, dc, B;                 // replicas, Doppler bins of this chunk, non-coherent blocks
  int Rc, G;                    // group shape: replicas x Doppler bins
  int ngr, ng;                  // replica groups per Doppler group; number of groups
  int nrt, ntiles, tpt;         // row tiles, column tiles, column tiles per ticket
  int nsets, slots_per_set;     // scratch ring: sets of Rc*G*B unit-block slots
  int nR, nC;                   // rows / columns tickets per group (fixed; tickets past a ragged edge are empty)
  int D, d0, n_lags, zmul;
  float scale;
};

struct FusedSync {
  int* ticket;                  // next ticket
  int* rows_done;               // [ng]
  int* cols_done;               // [ng]
  int* error;                   // set when a wait gave up
};

// Per-(replica, doppler, tile) partial result of the correlate kernel.
struct Part {
  unsigned long long key;   // (float bits of max q) << 32 | (0xffffffff - lag): max key = max q, ties -> lowest lag
  float sum;                // sum of q over the tile's lags
  float pad;
};

// Per-replica search result as it leaves the device (16 bytes, all-gather friendly).
struct Record {
  float metric;
  int lag;
  int dbin;      // index into the doppler list of this call, -1 if no metric was > 0
  int pad;
};

__device__ __forceinline__ unsigned long long pack_key(float v, int lag) {
  return ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xffffffffu - (unsigned)lag);
}

// Table NCO sample n of normalised frequency f (cycles/sample), phase 0:
// table[floor((f*n)*1024) mod 1024], with the reference's float64 rounding sequence
// (one rounded product, exact scaling by 1024, floor) — gnsstools/nco.py:7-9.
__device__ __forceinline__ float2 nco_sample(const float2* __restrict__ tab, double f, int n) {
  const double ph = __dmul_rn(f, (double)n);
  const long long k = __double2ll_rd(ph * 1024.0);
  return __ldg(&tab[(int)(k & (kNcoSize - 1))]);
}

template <int SRC>
__device__ __forceinline__ float2 load_input(const float2* __restrict__ x, const float* __restrict__ rep,
                                             const float2* __restrict__ nco_tab, double f, long long base, int n) {
  if (SRC == 0) {
    const float2 s = __ldg(&x[base + n]);
    return cmul(s, nco_sample(nco_tab, f, n));
  } else {
    return make_float2(__ldg(&rep[base + n]), 0.f);
  }
}

// Block-wide reduction of (key max, sum); result valid in thread 0.
__device__ __forceinline__ void block_reduce_part(unsigned long long& key, float& sum) {
  __shared__ unsigned long long s_key[32];
  __shared__ float s_sum[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, sum, o);
    key = k2 > key ? k2 : key;
    sum += s2;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();                         // protect s_key/s_sum reuse across calls
  if (lane == 0) { s_key[warp] = key; s_sum[warp] = sum; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    key = lane < nw ? s_key[lane] : 0ull;
    sum = lane < nw ? s_sum[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o);
      const float s2 = __shfl_xor_sync(0xffffffffu, sum, o);
      key = k2 > key ? k2 : key;
      sum += s2;
    }
  }
}

// ============================================================================ mid path
// Whole transform resident in one CTA. Shared memory: A[N1][SA] and Bt[N2][SB] float2
// (SA = N2|1, SB = N1|1 keep the transposes conflict-free), then optionally q[N] float.
__host__ __device__ __forceinline__ int mid_sa(int N2) { return N2 | 1; }
__host__ __device__ __forceinline__ int mid_sb(int N1) { return N1 | 1; }

// grid.x = number of transforms. SRC 0: transform t = d*B + b of the capture (block b,
// doppler d, wipe-off fused into the load). SRC 1: transform t = replica t (real input).
// Output: X[t*N + p2*N1 + p1], "position" order.
template <int RC, int SRC>
__global__ void __launch_bounds__(kThreads, 2)
k_fwd_mid(DevPlan pl, const float2* __restrict__ x, const float* __restrict__ rep,
          const double* __restrict__ freq, const float2* __restrict__ nco_tab,
          int stride, int B, float2* __restrict__ X) {
  GNSSACQ_DYN_SMEM(float2, sm);
  const int N = pl.N, N1 = pl.N1, N2 = pl.N2;
  const int SA = mid_sa(N2), SB = mid_sb(N1);
  float2* A = sm;
  float2* Bt = sm + N1 * SA;
  const int tc = threadIdx.x & (kTW - 1), tb = threadIdx.x / kTW, nb = blockDim.x / kTW;
  const int t = blockIdx.x;
  long long base;
  double f = 0.0;
  if (SRC == 0) { const int d = t / B, b = t - d * B; base = (long long)b * stride; f = freq[d]; }
  else { base = (long long)t * N; }

  for (int n1 = tb; n1 < N1; n1 += nb)
    for (int n2 = tc; n2 < N2; n2 += kTW)
      A[n1 * SA + n2] = (SRC == 0 && n1 * N2 + n2 >= pl.xlen) ? make_float2(0.f, 0.f) : load_input<SRC>(x, rep, nco_tab, f, base, n1 * N2 + n2);
  __syncthreads();
  subfft_tile<RC, false>(A, SA, N2, pl.s1);
  for (int p1 = tb; p1 < N1; p1 += nb)
    for (int n2 = tc; n2 < N2; n2 += kTW)
      Bt[n2 * SB + p1] = cmul(A[p1 * SA + n2], __ldg(&pl.twm[p1 * N2 + n2]));
  __syncthreads();
  subfft_tile<RC, false>(Bt, SB, N1, pl.s2);
  float2* out = X + (long long)t * N;
  for (int p2 = tb; p2 < N2; p2 += nb)
    for (int p1 = tc; p1 < N1; p1 += kTW)
      out[p2 * N1 + p1] = Bt[p2 * SB + p1];
}

// grid.x = R * Dc; CTA (r, dd) loops over the B non-coherent blocks.
// X: [Dc][B][N] capture spectra, C: [R][N] replica spectra (both position order).
template <int RC>
__global__ void __launch_bounds__(kThreads, 2)
k_corr_mid(DevPlan pl, const float2* __restrict__ X, const float2* __restrict__ C,
           int R, int B, int D, int d0, int n_lags, float scale,
           Part* __restrict__ parts, float* __restrict__ q_dump) {
  GNSSACQ_DYN_SMEM(float2, sm);
  const int N = pl.N, N1 = pl.N1, N2 = pl.N2;
  const int SA = mid_sa(N2), SB = mid_sb(N1);
  float2* A = sm;
  float2* Bt = sm + N1 * SA;
  float* qs = reinterpret_cast<float*>(Bt + N2 * SB);
  const int tc = threadIdx.x & (kTW - 1), tb = threadIdx.x / kTW, nb = blockDim.x / kTW;
  const int r = blockIdx.x % R, dd = blockIdx.x / R;
  const float2* Cr = C + (long long)r * N;
  unsigned long long key = 0ull;
  float sum = 0.f;

  for (int b = 0; b < B; ++b) {
    const float2* Xb = X + ((long long)dd * B + b) * N;
    for (int p2 = tb; p2 < N2; p2 += nb)
      for (int p1 = tc; p1 < N1; p1 += kTW)
        Bt[p2 * SB + p1] = cmulc(__ldg(&Cr[p2 * N1 + p1]), __ldg(&Xb[p2 * N1 + p1]));
    __syncthreads();
    subfft_tile<RC, true>(Bt, SB, N1, pl.s2);
    for (int p1 = tb; p1 < N1; p1 += nb)
      for (int n2 = tc; n2 < N2; n2 += kTW)
        A[p1 * SA + n2] = cmulc(Bt[n2 * SB + p1], __ldg(&pl.twm[p1 * N2 + n2]));
    __syncthreads();
    subfft_tile<RC, true>(A, SA, N2, pl.s1);
    const bool last = (b + 1 == B);
    for (int n1 = tb; n1 < N1; n1 += nb)
      for (int n2 = tc; n2 < N2; n2 += kTW) {
        const float2 v = A[n1 * SA + n2];
        const int lag = n1 * N2 + n2;
        float acc = __fsqrt_rn(v.x * v.x + v.y * v.y) * scale;
        if (b > 0) acc += qs[lag];
        if (!last) { qs[lag] = acc; }
        else {
          if (lag < pl.sum_lags) sum += acc;
          if (lag < n_lags) { const unsigned long long k = pack_key(acc, lag); key = k > key ? k : key; }
          if (q_dump) q_dump[((long long)r * D + d0 + dd) * N + lag] = acc;
        }
      }
    // the next iteration only writes Bt before its first barrier; A is rewritten after it.
  }
  block_reduce_part(key, sum);
  if (threadIdx.x == 0) {
    Part p; p.key = key; p.sum = sum; p.pad = 0.f;
    parts[(long long)r * D + d0 + dd] = p;
  }
}

// ============================================================================ large path
// N = N1*N2 too long for one CTA: a column kernel (length-N1 transforms at stride N2,
// 16 adjacent columns per CTA -> 128-byte global segments) and a row kernel (length-N2
// contiguous transforms, 16 rows per CTA, transposed into the tile with an odd pitch).
constexpr int kTileW = 16;
constexpr int kRowPitch = kTileW + 1;

// grid = (ceil(N2/16), transforms). Load (+wipe-off), column FFT, four-step twiddle, store.
template <int RC, int SRC>
__global__ void __launch_bounds__(kThreads, 2)
k_fwd_cols(DevPlan pl, const float2* __restrict__ x, const float* __restrict__ rep,
           const double* __restrict__ freq, const float2* __restrict__ nco_tab,
           int stride, int B, float2* __restrict__ X) {
  GNSSACQ_DYN_SMEM(float2, tile);
  const int N = pl.N, N1 = pl.N1, N2 = pl.N2;
  const int tc = threadIdx.x & (kTW - 1), tb = threadIdx.x / kTW, nb = blockDim.x / kTW;
  const int t = blockIdx.y;
  const int col0 = blockIdx.x * kTileW;
  const int ncols = imin(kTileW, N2 - col0);
  long long base;
  double f = 0.0;
  if (SRC == 0) { const int d = t / B, b = t - d * B; base = (long long)b * stride; f = freq[d]; }
  else { base = (long long)t * N; }
  if (tc < ncols)
    for (int n1 = tb; n1 < N1; n1 += nb) {
      const int n = n1 * N2 + col0 + tc;
      const int row = pl.gt ? __ldg(&pl.fpos1[n % N1]) : n1;
      tile[row * kTileW + tc] = (SRC == 0 && n >= pl.xlen) ? make_float2(0.f, 0.f) : load_input<SRC>(x, rep, nco_tab, f, base, n);
    }
  __syncthreads();
  subfft_tile<RC, false>(tile, kTileW, ncols, pl.s1);
  float2* out = X + (long long)t * N;
  if (tc < ncols)
    for (int p1 = tb; p1 < N1; p1 += nb) {
      const int g = p1 * N2 + col0 + tc;
      out[g] = pl.gt ? tile[p1 * kTileW + tc] : cmul(tile[p1 * kTileW + tc], __ldg(&pl.twm[g]));
    }
}

// grid = (ceil(N1/16), transforms). In-place row FFTs on X.
template <int RC>
__global__ void __launch_bounds__(kThreads, 2)
k_fwd_rows(DevPlan pl, float2* __restrict__ X) {
  GNSSACQ_DYN_SMEM(float2, tile);
  const int N = pl.N, N1 = pl.N1, N2 = pl.N2;
  const int tc = threadIdx.x & (kTW - 1), tb = threadIdx.x / kTW, nb = blockDim.x / kTW;
  const int row0 = blockIdx.x * kTileW;
  const int nrows = imin(kTileW, N1 - row0);
  float2* Xt = X + (long long)blockIdx.y * N;
  for (int c = tb; c < nrows; c += nb)
    for (int e = tc; e < N2; e += kTW)
      tile[(pl.gt ? __ldg(&pl.fpos2[e]) : e) * kRowPitch + c] = Xt[(row0 + c) * N2 + e];
  __syncthreads();
  subfft_tile<RC, false>(tile, kRowPitch, nrows, pl.s2);
  for (int c = tb; c < nrows; c += nb)
    for (int e = tc; e < N2; e += kTW)
      Xt[(row0 + c) * N2 + e] = tile[e * kRowPitch + c];
}

// grid = (ceil(N1/16), B, units). unit u -> replica r = (u0+u) % R, doppler dd = (u0+u) / R.
// Multiply by the replica spectrum, inverse row FFT, conjugate four-step twiddle -> scratch.
template <int RC>
__global__ void __launch_bounds__(kThreads, 2)
k_corr_rows(DevPlan pl, const float2* __restrict__ X, const float2* __restrict__ C,
            int R, int B, int u0, float2* __restrict__ scratch) {
  GNSSACQ_DYN_SMEM(float2, tile);
  const int N = pl.N, N1 = pl.N1, N2 = pl.N2;
  const int tc = threadIdx.x & (kTW - 1), tb = threadIdx.x / kTW, nb = blockDim.x / kTW;
  const int row0 = blockIdx.x * kTileW;
  const int nrows = imin(kTileW, N1 - row0);
  const int b = blockIdx.y, ul = blockIdx.z, u = u0 + ul;
  const int r = u % R, dd = u / R;
  const float2* Cr = C + (long long)r * N;
  const float2* Xb = X + ((long long)dd * B + b) * N;
  for (int c = tb; c < nrows; c += nb)
    for (int e = tc; e < N2; e += kTW) {
      const int g = (row0 + c) * N2 + e;
      tile[e * kRowPitch + c] = cmulc(__ldg(&Cr[g]), __ldg(&Xb[g]));
    }
  __syncthreads();
  subfft_tile<RC, true>(tile, kRowPitch, nrows, pl.s2);
  float2* out = scratch + ((long long)ul * B + b) * N;
  for (int c = tb; c < nrows; c += nb)
    for (int e = tc; e < N2; e += kTW) {
      const int g = (row0 + c) * N2 + e;
      out[g] = pl.gt ? tile[e * kRowPitch + c] : cmulc(tile[e * kRowPitch + c], __ldg(&pl.twm[g]));
    }
}

// grid = (units, ceil(N2/16)). Inverse column FFT of every block, |.|, non-coherent sum,
// per-tile max/argmax/sum. Shared memory: tile[N1][16] float2 (+ q[N1][16] float if B > 1).
template <int RC>
__global__ void __launch_bounds__(kThreads, 2)
k_corr_cols(DevPlan pl, const float2* __restrict__ scratch, int R, int B, int D, int d0, int u0,
            int n_lags, float scale, int ntiles, Part* __restrict__ parts, float* __restrict__ q_dump,
            unsigned* __restrict__ /*unit_hint: used by the specialised kernels only*/) {
  GNSSACQ_DYN_SMEM(float2, tile);
  const int N = pl.N, N1 = pl.N1, N2 = pl.N2;
  float* qs = reinterpret_cast<float*>(tile + N1 * kTileW);
  const int tc = threadIdx.x & (kTW - 1), tb = threadIdx.x / kTW, nb = blockDim.x / kTW;
  const int col0 = blockIdx.y * kTileW;                    // grid = (units, tiles), as k_corr_cols_s
  const int ncols = imin(kTileW, N2 - col0);
  const int ul = blockIdx.x, u = u0 + ul;
  const int r = u % R, dd = u / R;
  unsigned long long key = 0ull;
  float sum = 0.f;
  for (int b = 0; b < B; ++b) {
    const float2* in = scratch + ((long long)ul * B + b) * N;
    if (tc < ncols)
      for (int p1 = tb; p1 < N1; p1 += nb)
        tile[p1 * kTileW + tc] = in[p1 * N2 + col0 + tc];
    __syncthreads();
    subfft_tile<RC, true>(tile, kTileW, ncols, pl.s1);
    const bool last = (b + 1 == B);
    if (tc < ncols)
      for (int n1 = tb; n1 < N1; n1 += nb) {
        const float2 v = tile[n1 * kTileW + tc];
        int lag = n1 * N2 + __ldg(&pl.col_lag[col0 + tc]);
        if (lag >= N) lag -= N;
        float acc = __fsqrt_rn(v.x * v.x + v.y * v.y) * scale;
        if (b > 0) acc += qs[n1 * kTileW + tc];
        if (!last) { qs[n1 * kTileW + tc] = acc; }
        else {
          if (lag < pl.sum_lags) sum += acc;
          if (lag < n_lags) { const unsigned long long k = pack_key(acc, lag); key = k > key ? k : key; }
          if (q_dump) q_dump[((long long)r * D + d0 + dd) * N + lag] = acc;
        }
      }
    __syncthreads();     // tile is reloaded at the top of the next iteration
  }
  block_reduce_part(key, sum);
  if (threadIdx.x == 0) {
    Part p; p.key = key; p.sum = sum; p.pad = 0.f;
    parts[((long long)r * D + d0 + dd) * ntiles + blockIdx.y] = p;
  }
}

// ============================================================================ finalize
// grid = R, one CTA per replica: fold tiles, form the per-doppler metric and pick the best
// doppler bin with the reference's rule — strict '>' scanning ascending bins from 0, so ties
// go to the lowest bin and nothing is selected unless some metric is > 0
// (acquire-gps-l1.py:25,36-39).
// Doppler groups: the D entries of a call may be G consecutive groups of Dg (GLONASS FDMA: one
// group per channel, acquire-glonass-l1.py:26-39); grid = (R, G), record g*R + r, dbin relative to
// the group.
static __global__ void __launch_bounds__(128)
k_finalize(const Part* __restrict__ parts, int D, int Dg, int ntiles, int N, int normalize, Record* __restrict__ out) {
  const int r = blockIdx.x, g0 = blockIdx.y * Dg;
  unsigned long long best = 0ull;
  float dummy = 0.f;
  for (int d = threadIdx.x; d < Dg; d += blockDim.x) {
    const Part* p = parts + ((long long)r * D + g0 + d) * ntiles;
    unsigned long long key = 0ull;
    float sum = 0.f;
    for (int t = 0; t < ntiles; ++t) { key = p[t].key > key ? p[t].key : key; sum += p[t].sum; }
    const float peak = __uint_as_float((unsigned)(key >> 32));
    const float metric = normalize ? peak / (sum / (float)N) : peak;
    if (metric > 0.f) {
      const unsigned long long k = pack_key(metric, d);
      best = k > best ? k : best;
    }
  }
  block_reduce_part(best, dummy);
  if (threadIdx.x == 0) {
    Record rec; rec.metric = 0.f; rec.lag = 0; rec.dbin = -1; rec.pad = 0;
    if (best != 0ull) {
      const int d = (int)(0xffffffffu - (unsigned)(best & 0xffffffffull));
      const Part* p = parts + ((long long)r * D + g0 + d) * ntiles;
      unsigned long long key = 0ull;
      for (int t = 0; t < ntiles; ++t) key = p[t].key > key ? p[t].key : key;
      rec.metric = __uint_as_float((unsigned)(best >> 32));
      rec.lag = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
      rec.dbin = d;
    }
    out[(long long)blockIdx.y * gridDim.x + r] = rec;
  }
}

// Rank-ordered merge of Doppler-sharded records (all: [world][R], rank k owning the bins that start
// at off(k) = k*(D/world) + min(k, D%world)): a later rank replaces the best only on a strictly
// greater metric, so ties go to the lowest Doppler bin exactly as the reference's ascending scan
// (acquire-gps-l1.py:26,36-39). Identical on every rank.
static __global__ void __launch_bounds__(128)
k_merge_records(const Record* __restrict__ all, int world, int R, int D, Record* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int base = D / world, extra = D % world;
  Record best = all[r];
  if (best.dbin < 0) { best.metric = 0.f; best.lag = 0; best.dbin = -1; }
  for (int k = 1; k < world; ++k) {
    Record rec = all[(long long)k * R + r];
    if (rec.dbin >= 0 && rec.metric > best.metric) {
      rec.dbin += k * base + (k < extra ? k : extra);
      best = rec;
    }
  }
  best.pad = 0;
  out[r] = best;
}

// int8 replicas (+-1 / 0) -> float32: lets callers ship a quarter of the bytes over PCIe.
static __global__ void __launch_bounds__(kThreads)
k_i8_to_f32(const signed char* __restrict__ in, long long n, float* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (float)in[i];
}

// Periodic extension of a length-n replica into a length-M transform (M >= 2n-1): sample m < 2n-1
// takes rep[m mod n], the rest is zero, so that the circular correlation of length M equals the
// circular correlation of length n at lags 0..n-1 when the capture block is zero-padded.
static __global__ void __launch_bounds__(kThreads)
k_extend_replicas(const float* __restrict__ rep, int n, int M, float* __restrict__ out) {
  const long long r = blockIdx.y;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x)
    out[r * M + m] = m < 2 * n - 1 ? rep[r * n + (m >= n ? m - n : m)] : 0.f;
}

// ============================================================================ capture mix
// In-place whole-capture carrier wipe-off, reference gnsstools/nco.py:30-41: int64 phase
// accumulator scaled by 2^50 (closed form dp_i = dp0 + i*df, wrapping), complex128 table,
// complex128 product rounded to complex64 on store.
static __global__ void __launch_bounds__(kThreads, 2)
k_mix(float2* __restrict__ x, long long n, unsigned long long dp0, unsigned long long df,
      const double2* __restrict__ tab) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long dp = (long long)(dp0 + (unsigned long long)i * df);
    const double2 w = tab[(int)((dp >> 50) & (kNcoSize - 1))];
    const float2 s = x[i];
    const double sr = (double)s.x, si = (double)s.y;
    const double re = __dadd_rn(__dmul_rn(sr, w.x), -__dmul_rn(si, w.y));
    const double im = __dadd_rn(__dmul_rn(sr, w.y), __dmul_rn(si, w.x));
    x[i] = make_float2((float)re, (float)im);
  }
}

}  // namespace acq
