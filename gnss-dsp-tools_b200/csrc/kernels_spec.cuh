// Plan-specialised correlate kernels for the transform lengths GNSS receivers actually use.
//
// Same algorithm and data layout as the runtime-planned k_corr_rows / k_corr_cols
// (kernels.cuh), but the radix schedule of the tile transform is a template parameter, so
// strides, trip counts and twiddle offsets are immediates, and the first / last butterfly
// stages are fused with the global-memory traffic around them:
//   rows kernel: [load X, C -> multiply -> first inverse stage] -> smem stages ->
//                [last inverse stage -> conjugate four-step twiddle -> coalesced store]
//   cols kernel: [coalesced load -> first inverse stage] -> smem stages ->
//                [last inverse stage -> |.| -> non-coherent sum -> max/argmax/sum]
// Lengths without a specialisation run the generic kernels; results agree to rounding.
#pragma once
#include "kernels.cuh"

namespace acq {

// Radix schedule of one tile transform, forward order (as fft_plan.h builds it).
template <int F_, int R0_, int R1_ = 1, int R2_ = 1, int R3_ = 1>
struct Sub {
  static constexpr int F = F_;
  static constexpr int NS = R1_ == 1 ? 1 : (R2_ == 1 ? 2 : (R3_ == 1 ? 3 : 4));
  static_assert(R0_ * R1_ * R2_ * R3_ == F_, "radices must multiply to F");
  static_assert(NS >= 2, "specialised kernels need at least two stages");
  __host__ __device__ static constexpr int radix(int j) { return j == 0 ? R0_ : (j == 1 ? R1_ : (j == 2 ? R2_ : R3_)); }
  __host__ __device__ static constexpr int stride(int j) {      // product of the later radices
    int m = 1;
    for (int k = j + 1; k < 4; ++k) m *= radix(k);
    return m;
  }
  __host__ __device__ static constexpr int gcd_(int a, int b) { return b == 0 ? a : gcd_(b, a % b); }
  __host__ __device__ static constexpr bool coprime_() {
    for (int j = 0; j < NS; ++j)
      for (int k = j + 1; k < NS; ++k)
        if (gcd_(radix(j), radix(k)) != 1) return false;
    return true;
  }
  // Pairwise-coprime radices: the kernels run the transform in its prime-factor form — a plain
  // multi-dimensional DFT over the digits of the tile position, no stage twiddles; the host plan
  // (fft_plan.h::set_pfa) supplies the index maps that go with it.
#ifdef GNSSACQ_NO_PFA
  static constexpr bool kPfa = false;       // A/B builds of tools/microbench only
#else
  static constexpr bool kPfa = coprime_();
#endif
};

constexpr bool is_split_radix(int R) { return R == 31; }
#ifdef GNSSACQ_NO_FWD_RADER31
constexpr bool kFwdRader31 = false;      // A/B builds: the forward radix-31 stage shared by a warp pair
#else
constexpr bool kFwdRader31 = true;
#endif
// butterflies a thread keeps in flight per loop trip: small radices need several for ILP
constexpr int stage_unroll(int R) { return R <= 5 ? 4 : (R <= 10 ? 2 : 1); }
// butterflies whose global loads the columns kernel's first stage keeps in flight per thread
#ifdef GNSSACQ_COLS_LOAD_UNROLL
constexpr int cols_load_unroll(int) { return GNSSACQ_COLS_LOAD_UNROLL; }      // A/B builds of tools/microbench only
#else
constexpr int cols_load_unroll(int R) { return stage_unroll(R); }
#endif

// ---- inverse butterfly on registers: v[q] (already conj-twiddled) -> natural-order outputs.
// IDFT(x)[k] = DFT(x)[-k mod R]: the forward butterfly followed by an index reversal, which is pure
// register renaming in unrolled code. (The re/im swap form costs two MOVs per element once the
// values live in the 64-bit register pairs the packed f32x2 instructions need: ncu counted 70
// MOVs per radix-32 butterfly.)
template <int R> __device__ __forceinline__ void inv_dft(float2* v) {
  Dft<R>::run(v);
#pragma unroll
  for (int q = 1; q < (R + 1) / 2; ++q) { const float2 t = v[q]; v[q] = v[R - q]; v[R - q] = t; }
}

// ---- one in-shared-memory inverse stage (stage J of S). Element e of column c lives at
// tile[e*ES + c*CS]: the columns kernel uses (ES, CS) = (16, 1), the rows kernel (1, odd pitch)
// — either way the 16 lanes of a half-warp (consecutive c) hit 16 distinct bank pairs.
template <class S, int J, int ES, int CS, int THREADS = kThreads>
__device__ __forceinline__ void inv_stage_smem(float2* tile, int ncols, const float2* __restrict__ twbase, int twoff) {
  constexpr int R = S::radix(J), m = S::stride(J), nbf = S::F / R;
  const int tc = threadIdx.x & (kTW - 1);
  if constexpr (is_split_radix(R)) {
    stage_tile_split<R, true, ES, CS, S::kPfa>(tile, ncols, S::F, m, twbase);   // warp-pair version, W_F table
  } else {
    const int tb = threadIdx.x / kTW;
    constexpr int nb = THREADS / kTW;
    const float2* tws = twbase + twoff;
    // When the butterfly stride divides the 16 butterfly groups, a thread meets the same
    // twiddle row in every iteration (i = tb mod m): load it once.
    constexpr bool kHoist = !S::kPfa && m > 1 && nb % m == 0 && R <= 8;
    float2 wh[kHoist ? R : 1];
    if constexpr (kHoist) {
      const float2* w = tws + (tb % m) * (R - 1);
#pragma unroll
      for (int q = 1; q < R; ++q) wh[q] = __ldg(&w[q - 1]);
    }
    if (tc < ncols) {
#pragma unroll stage_unroll(R)
      for (int bf = tb; bf < nbf; bf += nb) {
        const int blk = bf / m, i = bf - blk * m;
        float2* p = tile + (blk * R * m + i) * ES + tc * CS;
        float2 v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = p[q * m * ES];
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
        for (int q = 0; q < R; ++q) p[q * m * ES] = v[q];
      }
    }
  }
  __syncthreads();
}

template <class S, int J, int JEND, int ES, int CS, int THREADS = kThreads>
__device__ __forceinline__ void inv_stages_smem(float2* tile, int ncols, const SubPlan& sp) {
  if constexpr (J >= JEND) {
    inv_stage_smem<S, J, ES, CS, THREADS>(tile, ncols, sp.tw, sp.tws_off[J]);
    inv_stages_smem<S, J - 1, JEND, ES, CS, THREADS>(tile, ncols, sp);
  }
}

// =========================================================================== rows kernel
// grid = (ceil(N1/16), B, units), as k_corr_rows. S = schedule of the length-N2 transform.
// The tile is row-major like memory, tile[c*P + e] with an odd pitch P, so the global loads
// and stores are straight coalesced row copies and no transposition is needed.
template <class S> __host__ __device__ constexpr int rows_pitch() { return S::F | 1; }
template <class S> __host__ __device__ constexpr size_t rows_spec_smem() { return (size_t)kTileW * rows_pitch<S>() * sizeof(float2); }

template <class S> __host__ __device__ constexpr int rows_min_ctas() { return rows_spec_smem<S>() * 3 <= 200 * 1024 ? 3 : 2; }

template <class S, bool GT = false>
__global__ void __launch_bounds__(kThreads, rows_min_ctas<S>())
k_corr_rows_s(DevPlan pl, const float2* __restrict__ X, const float2* __restrict__ C,
              int R_, int B, int u0, float2* __restrict__ scratch) {
  GNSSACQ_DYN_SMEM(float2, tile);
  constexpr int N2 = S::F, P = rows_pitch<S>(), NS = S::NS;
  const int N = pl.N, N1 = pl.N1;
  const int tc = threadIdx.x & (kTW - 1), tb = threadIdx.x / kTW;
  constexpr int nb = kThreads / kTW;
  const int row0 = blockIdx.x * kTileW;
  const int nrows = imin(kTileW, N1 - row0);
  const int b = blockIdx.y, ul = blockIdx.z, u = u0 + ul;
  const int r = u % R_, dd = u / R_;
  const float2* Cr = C + (long long)r * N + (long long)row0 * N2;
  const float2* Xb = X + ((long long)dd * B + b) * N + (long long)row0 * N2;

  // ---- coalesced load of the 16 rows, multiplied by the replica spectrum on the way in
  for (int c = tb; c < nrows; c += nb) {
    const float2* xr = Xb + c * N2;
    const float2* cr = Cr + c * N2;
    float2* trow = tile + c * P;
#pragma unroll 8
    for (int e = tc; e < N2; e += kTW) trow[e] = cmulc(__ldg(&cr[e]), __ldg(&xr[e]));
  }
  __syncthreads();
  // ---- inverse stages NS-1 .. 1 in shared memory
  inv_stages_smem<S, NS - 1, 1, 1, P>(tile, nrows, pl.s2);
  // ---- last inverse stage (first forward stage, stride m0)
  float2* out = scratch + ((long long)ul * B + b) * N + (long long)row0 * N2;
  const float2* twm = pl.twm_inv + (long long)row0 * N2;
  constexpr int R0 = S::radix(0), m0 = S::stride(0);
  if constexpr (!is_split_radix(R0) && m0 >= 16) {
    // fused with the conjugate four-step twiddle and the store: lanes walk i (consecutive n2),
    // stage twiddles come from the transposed table (q-major) so lanes read them contiguously
    const float2* twt = pl.s2.tw + pl.s2.tws0_t_off;
    const int items = m0 * nrows;
    for (int id = threadIdx.x; id < items; id += kThreads) {
      const int c = id / m0, i = id - c * m0;
      const float2* p = tile + c * P + i;
      float2 v[R0];
#pragma unroll
      for (int q = 0; q < R0; ++q) v[q] = p[q * m0];
      if constexpr (!S::kPfa) {
#pragma unroll
        for (int q = 1; q < R0; ++q) v[q] = cmulc(v[q], __ldg(&twt[(q - 1) * m0 + i]));
      }
      inv_dft<R0>(v);
      const int g = c * N2 + i;
#pragma unroll
      for (int q = 0; q < R0; ++q) out[g + q * m0] = GT ? v[q] : cmulc(v[q], __ldg(&twm[g + q * m0]));
    }
  } else {
    inv_stage_smem<S, 0, 1, P>(tile, nrows, pl.s2.tw, pl.s2.tws_off[0]);
    for (int c = tb; c < nrows; c += nb)
      for (int e = tc; e < N2; e += kTW) {
        const int g = c * N2 + e;
        out[g] = GT ? tile[c * P + e] : cmulc(tile[c * P + e], __ldg(&twm[g]));
      }
  }
}

// =========================================================================== cols kernel
// grid = (units, ceil(N2/16)): units fastest, so the tiles of one unit are spread over the waves and
// all but the first start with a warm peak-search floor (unit_hint, see kernels_v3.cuh).
// S = schedule of the length-N1 transform. MULTI = more than one non-coherent block (q kept in shared
// memory between blocks).
__device__ __forceinline__ float sqrt_fast(float a) {
#if defined(__CUDA_ARCH__)
  float r;
  // one MUFU.SQRT, <= 1 ulp: far inside the 1e-4 budget. ftz: without it ptxas wraps the MUFU in a
  // denormal rescue (FSETP + 2 FMUL per output, ncu: 93 extra FMULs per radix-31 butterfly); inputs are
  // |v|^2 of correlation sums, never subnormal unless the capture is all zeros (then 0 either way).
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
#else
  return sqrtf(a);
#endif
}

// Last inverse stage of the columns transform fused with |.|, the non-coherent sum and the
// running peak search (shared by the one-tile-per-CTA and the pipelined columns kernels).
// Per-thread state: best / bestlag / sum over the eligible lags, unscaled.
// SPLIT: share a radix-31 butterfly between two warps (halves its live registers; needed at 256
// threads x 3 CTAs per SM). With 128-thread CTAs the butterfly fits in one thread's 128 registers
// and the duplicated loads / twiddle multiplies of the shared form go away.
// TW = tile width (columns side by side in the tile, a power of two <= 16).
template <class S, bool MULTI, int THREADS = kThreads, bool SPLIT = true, int TW = kTW>
__device__ __forceinline__ void cols_last_stage(const float2* tile, float* qs, const DevPlan& pl, int ncols, int lag0,
                                                int b, bool last, int n_lags, float scale, float* qd,
                                                float& best, int& bestlag, float& sum, float hint = 0.f) {
  // hint: the best eligible value an earlier tile of the same (replica, Doppler) unit reported (0 if none):
  // a lower bound of the unit maximum, so candidates below it cannot be the unit's peak (kernels_v3.cuh).
  static_assert(!SPLIT || TW == kTW, "the warp-pair butterfly assumes 16-column tiles");
  constexpr int WP = TW;
  const int N2 = pl.N2;
  const int tc = threadIdx.x & (TW - 1), tb = threadIdx.x / TW;
  constexpr int nb = THREADS / TW;
  const bool dump = qd != nullptr;
  constexpr int R0 = S::radix(0), m0 = S::stride(0);
  // A butterfly output sits at tile position pos = i + q*m0; the lag it stands for is
  // n1 * N2 + n2 with the time indices of that position and of this thread's column. Prime-factor
  // plans: n1 = (n1_of_pos[i] + q * F/R0) mod F (one table read per butterfly, outside the
  // peak-search branch: with ~46 outputs per thread and tile, some lane of a warp takes that
  // branch for most outputs, so its body must stay a few integer instructions);
  // Cooley-Tukey plans: n1 = pos; n2 = n2_of_pos[column] (the identity unless the rows transform is prime-factor).
  // the column map belongs to the length-N2 transform, whatever S (the length-N1 schedule) is: always through the table
  const int lagc = tc < ncols ? __ldg(&pl.col_lag[lag0]) : 0;       // n2 of the column; N1 * n2 for a coprime split
  const int Nfull = pl.N;
  auto lag_of = [&](int n1) -> int { const int l = n1 * N2 + lagc; return l >= Nfull ? l - Nfull : l; };   // wraps only when pl.gt
  auto n1_base = [&](int i) -> int { if constexpr (S::kPfa) return __ldg(&pl.n1_of_pos[i]); else return i; };
  auto n1_of = [&](int n1i, int q) -> int {
    if constexpr (S::kPfa) { const int n = n1i + q * (S::F / R0); return n >= S::F ? n - S::F : n; }
    else return n1i + q * m0;
  };
  // Peak search over a batch of NB butterfly outputs v[t] (output index q = qof(t), tile position
  // i + q*m0). Per output the common path is |v|, the sum and one max; one branch per batch asks
  // whether anything in it reaches the threshold — this thread's best, raised to the best any
  // converged lane of the warp holds (a hint that never exceeds the tile maximum, so no candidate
  // for the tile's (max, lowest lag) is skipped). Per-output branches serialise the |.| latency
  // chains and, with ~46 outputs per thread and tile, some lane of a warp takes them for most
  // outputs; per batch with the warp-wide threshold the slow path runs for a few batches per tile.
  auto sink_batch = [&](auto NBc, auto qof, const float2* v, int i, int n1i) {
    constexpr int NB = decltype(NBc)::value;
    float acc[NB];
#pragma unroll
    for (int t = 0; t < NB; ++t) acc[t] = sqrt_fast(v[t].x * v[t].x + v[t].y * v[t].y);
    if (MULTI) {
#pragma unroll
      for (int t = 0; t < NB; ++t) {
        const int pos = i + qof(t) * m0;
        if (b > 0) acc[t] += qs[pos * WP + tc];
        if (!last) qs[pos * WP + tc] = acc[t];
      }
      if (!last) return;
    }
    float m = acc[0];
#pragma unroll
    for (int t = 0; t < NB; ++t) { sum += acc[t]; m = fmaxf(m, acc[t]); }
#if defined(__CUDA_ARCH__)
    const float flo = fmaxf(best, hint);
    const float thr = fmaxf(flo, __uint_as_float(__reduce_max_sync(__activemask(), __float_as_uint(fmaxf(flo, 0.f)))));
#else
    const float flo = fmaxf(best, hint);
    const float thr = flo;
#endif
    if (m >= thr) {
#pragma unroll
      for (int t = 0; t < NB; ++t) {
        if (acc[t] >= flo) {
          const int lag = lag_of(n1_of(n1i, qof(t)));
          if (lag < n_lags && (acc[t] > best || (acc[t] == best && lag < bestlag))) { best = acc[t]; bestlag = lag; }
        }
      }
    }
    if (dump) {
#pragma unroll
      for (int t = 0; t < NB; ++t) qd[lag_of(n1_of(n1i, qof(t)))] = acc[t] * scale;
    }
  };
  // batch handler for prime_outputs_batched: pairs k0..k0+nk-1 -> outputs k and R0-k (|.| ignores the re/im swap)
  auto prime_batch = [&](int i, int n1i) {
    return [&, i, n1i](auto K0c, auto NKc, const float2* re, const float2* im) {
      constexpr int k0 = decltype(K0c)::value, nk = decltype(NKc)::value;
      float2 v[2 * nk];
#pragma unroll
      for (int t = 0; t < nk; ++t) {
        v[2 * t] = make_float2(re[t].x + im[t].y, re[t].y - im[t].x);
        v[2 * t + 1] = make_float2(re[t].x - im[t].y, re[t].y + im[t].x);
      }
      sink_batch(std::integral_constant<int, 2 * nk>{}, [](int t) { return (t & 1) ? R0 - (k0 + t / 2) : k0 + t / 2; }, v, i, n1i);
    };
  };
  auto sink_one = [&](int q, float2 v, int i, int n1i) {
    sink_batch(std::integral_constant<int, 1>{}, [q](int) { return q; }, &v, i, n1i);
  };
  // leg q of the butterfly at tile offset p, conjugate stage twiddle applied unless prime-factor
  auto leg = [&](const float2* p, const float2* w, int q) -> float2 {
    float2 v = p[q * m0 * WP];
    if constexpr (!S::kPfa) v = cmulc(v, __ldg(&w[q - 1]));
    return v;
  };
  if constexpr (is_split_radix(R0) && SPLIT) {
    // warp-pair butterfly (see stage_tile_split): both warps load, each emits half the outputs
    constexpr int H = (R0 - 1) / 2, KA = (H + 1) / 2;
    const int warp = threadIdx.x >> 5, role = warp & 1;
    const int slot = (warp >> 1) * 2 + ((threadIdx.x >> 4) & 1);
    constexpr int nslots = THREADS >> 5;
    const float2* tws = pl.s1.tw + pl.s1.tws_off[0];
    for (int i = slot; i < m0; i += nslots) {
      if (tc < ncols) {
        const float2* p = tile + i * WP + tc;
        const float2* w = tws + i * (R0 - 1);
        float2 a[H + 1], bq[H + 1];
        const float2 x0 = cswap(p[0]);
        static_for<1, H + 1>([&](auto J) {
          constexpr int j = decltype(J)::value;
          const float2 s = cswap(leg(p, w, j));
          const float2 t = cswap(leg(p, w, R0 - j));
          a[j] = cadd(s, t);
          bq[j] = csub(s, t);
        });
        const int n1i = n1_base(i);
        if (role == 0) {
          float2 s0 = x0;
          static_for<1, H + 1>([&](auto J) { s0 = cadd(s0, a[decltype(J)::value]); });
          sink_one(0, s0, i, n1i);
          prime_outputs_batched<R0, 1, KA>(x0, a, bq, prime_batch(i, n1i));
        } else {
          prime_outputs_batched<R0, KA + 1, H>(x0, a, bq, prime_batch(i, n1i));
        }
      }
    }
  } else if constexpr (R0 >= 11 && R0 % 2 == 1) {
    // large prime, one thread per butterfly: outputs go straight into the sink, never into registers
    constexpr int H = (R0 - 1) / 2;
    const float2* tws = pl.s1.tw + pl.s1.tws_off[0];
    if (tc < ncols) {
      for (int i = tb; i < m0; i += nb) {
        const float2* p = tile + i * WP + tc;
        const float2* w = tws + i * (R0 - 1);
        float2 a[H + 1], bq[H + 1];
        const float2 x0 = cswap(p[0]);
        float2 s0 = x0;
        static_for<1, H + 1>([&](auto J) {
          constexpr int j = decltype(J)::value;
          const float2 s = cswap(leg(p, w, j));
          const float2 t = cswap(leg(p, w, R0 - j));
          a[j] = cadd(s, t);
          bq[j] = csub(s, t);
          s0 = cadd(s0, a[j]);
        });
        const int n1i = n1_base(i);
        sink_one(0, s0, i, n1i);
        prime_outputs_batched<R0, 1, H>(x0, a, bq, prime_batch(i, n1i));
      }
    }
  } else {
    const float2* tws = pl.s1.tw + pl.s1.tws_off[0];
    if (tc < ncols) {
#pragma unroll stage_unroll(R0)
      for (int i = tb; i < m0; i += nb) {
        const float2* p = tile + i * WP + tc;
        float2 v[R0];
#pragma unroll
        for (int q = 0; q < R0; ++q) v[q] = p[q * m0 * WP];
        if constexpr (!S::kPfa) {
          const float2* w = tws + i * (R0 - 1);
#pragma unroll
          for (int q = 1; q < R0; ++q) v[q] = cmulc(v[q], __ldg(&w[q - 1]));
        }
#pragma unroll
        for (int q = 0; q < R0; ++q) v[q] = cswap(v[q]);
        Dft<R0>::run(v);                                                 // |.| ignores the swap back
        sink_batch(std::integral_constant<int, R0>{}, [](int t) { return t; }, v, i, n1_base(i));
      }
    }
  }
}

// Three CTAs per SM (<= 85 registers) pay off except for the radix-16 schedules, whose
// butterflies need the registers (measured: 372 = 31*3*4 gains 8 %, 256 = 16*16 loses 12 %).
template <class S> __host__ __device__ constexpr int cols_min_ctas() {
  for (int j = 0; j < S::NS; ++j)
    if (S::radix(j) == 16) return 2;
  return 3;
}
template <class S, bool MULTI, int THREADS = kThreads, int MINCTAS = cols_min_ctas<S>(), bool SPLIT = true>
__global__ void __launch_bounds__(THREADS, MINCTAS)
k_corr_cols_s(DevPlan pl, const float2* __restrict__ scratch, int R_, int B, int D, int d0, int u0,
              int n_lags, float scale, int ntiles, Part* __restrict__ parts, float* __restrict__ q_dump,
              unsigned* __restrict__ unit_hint) {
  GNSSACQ_DYN_SMEM(float2, tile);
  constexpr int N1 = S::F, WP = kTileW, NS = S::NS;
  const int N = pl.N, N2 = pl.N2;
  float* qs = reinterpret_cast<float*>(tile + N1 * WP);
  const int tc = threadIdx.x & (kTW - 1), tb = threadIdx.x / kTW;
  constexpr int nb = THREADS / kTW;
  const int tileno = blockIdx.y;
  const int col0 = tileno * kTileW;
  const int ncols = imin(kTileW, N2 - col0);
  const int ul = blockIdx.x, u = u0 + ul;
  const int r = u % R_, dd = u / R_;
  const long long unit = (long long)r * D + d0 + dd;
  // Per-thread running peak over the eligible lags (unscaled; 1/N is applied once at the end).
  float best = -1.f, sum = 0.f;
  int bestlag = 0x7fffffff;
  const float hint = unit_hint ? __uint_as_float(__ldcg(&unit_hint[unit])) : 0.f;
  float* qd = q_dump ? q_dump + unit * N : nullptr;
  const int lag0 = col0 + tc;

  for (int b = 0; b < B; ++b) {
    const float2* in = scratch + ((long long)ul * B + b) * N + lag0;
    const bool last = (b + 1 == B);
    // ---- first inverse stage fused with the coalesced load
    {
      constexpr int R = S::radix(NS - 1), nbf = N1 / R;
      if (tc < ncols) {
#pragma unroll (cols_load_unroll(R))
        for (int bf = tb; bf < nbf; bf += nb) {
          float2 v[R];
#pragma unroll
          for (int q = 0; q < R; ++q) v[q] = in[(bf * R + q) * N2];
          inv_dft<R>(v);
#pragma unroll
          for (int q = 0; q < R; ++q) tile[(bf * R + q) * WP + tc] = v[q];
        }
      }
      __syncthreads();
    }
    inv_stages_smem<S, NS - 2, 1, WP, 1, THREADS>(tile, ncols, pl.s1);
    cols_last_stage<S, MULTI, THREADS, SPLIT>(tile, qs, pl, ncols, lag0, b, last, n_lags, scale, qd, best, bestlag, sum, hint);
    if (MULTI && !last) __syncthreads();       // tile and q are reused by the next block
  }
  // reduce on the unscaled value (the hint must be bit-exactly a value that occurred; scaling by 1/N is monotonic)
  unsigned long long key = bestlag != 0x7fffffff ? pack_key(best, bestlag) : 0ull;
  sum *= scale;
  block_reduce_part(key, sum);
  if (threadIdx.x == 0) {
    Part p; p.key = 0ull; p.sum = sum; p.pad = 0.f;
    if (key != 0ull) {
      const unsigned vb = (unsigned)(key >> 32);
      p.key = ((unsigned long long)__float_as_uint(__uint_as_float(vb) * scale) << 32) | (key & 0xffffffffull);
      if (unit_hint) atomicMax(&unit_hint[unit], vb);
    }
    parts[unit * ntiles + tileno] = p;
  }
}

// =========================================================================== forward kernels
// Same two-kernel four-step as k_fwd_cols / k_fwd_rows with the schedule as a template
// parameter. Forward stages are decimation-in-frequency: butterfly, then twiddle the outputs.
template <class S, int J, int ES, int CS>
__device__ __forceinline__ void fwd_stage_smem(float2* tile, int ncols, const float2* __restrict__ twbase, int twoff) {
  constexpr int R = S::radix(J), m = S::stride(J), nbf = S::F / R;
  const int tc = threadIdx.x & (kTW - 1);
  if constexpr (R == 31 && S::kPfa && kRader31On && kFwdRader31) {
    // one thread per butterfly, radix 31 as two 15-point convolutions (fft_core.cuh); a butterfly's 31 elements are its own
    const int tb = threadIdx.x / kTW;
    constexpr int nb = kThreads / kTW;
    if (tc < ncols)
      for (int bf = tb; bf < nbf; bf += nb) {
        const int blk = bf / m, i = bf - blk * m;
        float2* p = tile + (blk * R * m + i) * ES + tc * CS;
        float2 v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = p[q * m * ES];
        Dft31Rader::run(v);
#pragma unroll
        for (int q = 0; q < R; ++q) p[q * m * ES] = v[q];
      }
  } else if constexpr (is_split_radix(R)) {
    stage_tile_split<R, false, ES, CS, S::kPfa>(tile, ncols, S::F, m, twbase);
  } else {
    const int tb = threadIdx.x / kTW;
    constexpr int nb = kThreads / kTW;
    const float2* tws = twbase + twoff;
    if (tc < ncols) {
#pragma unroll stage_unroll(R)
      for (int bf = tb; bf < nbf; bf += nb) {
        const int blk = bf / m, i = bf - blk * m;
        float2* p = tile + (blk * R * m + i) * ES + tc * CS;
        float2 v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = p[q * m * ES];
        Dft<R>::run(v);
        if constexpr (m > 1 && !S::kPfa) {
          const float2* w = tws + i * (R - 1);
#pragma unroll
          for (int q = 1; q < R; ++q) v[q] = cmul(v[q], __ldg(&w[q - 1]));
        }
#pragma unroll
        for (int q = 0; q < R; ++q) p[q * m * ES] = v[q];
      }
    }
  }
  __syncthreads();
}

template <class S, int J, int JEND, int ES, int CS>
__device__ __forceinline__ void fwd_stages_smem(float2* tile, int ncols, const SubPlan& sp) {
  if constexpr (J <= JEND) {
    fwd_stage_smem<S, J, ES, CS>(tile, ncols, sp.tw, sp.tws_off[J]);
    fwd_stages_smem<S, J + 1, JEND, ES, CS>(tile, ncols, sp);
  }
}

// grid = (ceil(N2/16), transforms), as k_fwd_cols. S = schedule of the length-N1 transform.
template <class S, int SRC>
__global__ void __launch_bounds__(kThreads, 2)
k_fwd_cols_s(DevPlan pl, const float2* __restrict__ x, const float* __restrict__ rep,
             const double* __restrict__ freq, const float2* __restrict__ nco_tab,
             int stride, int B, float2* __restrict__ X) {
  GNSSACQ_DYN_SMEM(float2, tile);
  constexpr int N1 = S::F, WP = kTileW, NS = S::NS;
  const int N = pl.N, N2 = pl.N2;
  const int tc = threadIdx.x & (kTW - 1), tb = threadIdx.x / kTW;
  constexpr int nb = kThreads / kTW;
  const int t = blockIdx.y;
  const int col0 = blockIdx.x * kTileW;
  const int ncols = imin(kTileW, N2 - col0);
  long long base;
  double f = 0.0;
  if (SRC == 0) { const int d = t / B, b = t - d * B; base = (long long)b * stride; f = freq[d]; }
  else { base = (long long)t * N; }
  constexpr int R0 = S::radix(0), m0 = S::stride(0);
  // tile position p1 takes time sample n1_of_pos[p1] (identity unless the plan is prime-factor)
  auto n1_at = [&](int p1) -> int { if constexpr (S::kPfa) return __ldg(&pl.n1_of_pos[p1]); else return p1; };
  if (pl.gt) {
    // coprime split: rows of memory are m = n div N2; sample n = m*N2 + b of column b belongs to
    // the tile position fpos1[n mod N1] (a different row rotation per column, loads stay coalesced)
    if (tc < ncols)
      for (int m = tb; m < N1; m += nb) {
        const int n = m * N2 + col0 + tc;
        tile[__ldg(&pl.fpos1[n % N1]) * WP + tc] = load_input<SRC>(x, rep, nco_tab, f, base, n);
      }
    __syncthreads();
    fwd_stage_smem<S, 0, WP, 1>(tile, ncols, pl.s1.tw, pl.s1.tws_off[0]);
  } else if constexpr (is_split_radix(R0)) {
    if (tc < ncols)
      for (int n1 = tb; n1 < N1; n1 += nb)
        tile[n1 * WP + tc] = load_input<SRC>(x, rep, nco_tab, f, base, n1_at(n1) * N2 + col0 + tc);
    __syncthreads();
    fwd_stage_smem<S, 0, WP, 1>(tile, ncols, pl.s1.tw, pl.s1.tws_off[0]);
  } else {
    // first stage fused with the load and the carrier wipe-off
    const float2* tws = pl.s1.tw + pl.s1.tws_off[0];
    if (tc < ncols) {
#pragma unroll stage_unroll(R0)
      for (int i = tb; i < m0; i += nb) {
        float2 v[R0];
#pragma unroll
        for (int q = 0; q < R0; ++q) v[q] = load_input<SRC>(x, rep, nco_tab, f, base, n1_at(i + q * m0) * N2 + col0 + tc);
        Dft<R0>::run(v);
        if constexpr (!S::kPfa) {
          const float2* w = tws + i * (R0 - 1);
#pragma unroll
          for (int q = 1; q < R0; ++q) v[q] = cmul(v[q], __ldg(&w[q - 1]));
        }
#pragma unroll
        for (int q = 0; q < R0; ++q) tile[(i + q * m0) * WP + tc] = v[q];
      }
    }
    __syncthreads();
  }
  fwd_stages_smem<S, 1, NS - 2, WP, 1>(tile, ncols, pl.s1);
  // last stage (unit stride, no stage twiddle) fused with the four-step twiddle and the store
  {
    constexpr int R = S::radix(NS - 1), nbf = N1 / R;
    float2* out = X + (long long)t * N + col0 + tc;
    const float2* twm = pl.twm + col0 + tc;
    const bool gt = pl.gt != 0;
    if (tc < ncols) {
#pragma unroll stage_unroll(R)
      for (int bf = tb; bf < nbf; bf += nb) {
        float2 v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = tile[(bf * R + q) * WP + tc];
        Dft<R>::run(v);
#pragma unroll
        for (int q = 0; q < R; ++q) out[(bf * R + q) * N2] = gt ? v[q] : cmul(v[q], __ldg(&twm[(bf * R + q) * N2]));
      }
    }
  }
}

// grid = (ceil(N1/16), transforms), as k_fwd_rows: in-place length-N2 row transforms.
template <class S>
__global__ void __launch_bounds__(kThreads, rows_min_ctas<S>())
k_fwd_rows_s(DevPlan pl, float2* __restrict__ X) {
  GNSSACQ_DYN_SMEM(float2, tile);
  constexpr int N2 = S::F, P = rows_pitch<S>(), NS = S::NS;
  const int N = pl.N, N1 = pl.N1;
  const int tc = threadIdx.x & (kTW - 1), tb = threadIdx.x / kTW;
  constexpr int nb = kThreads / kTW;
  const int row0 = blockIdx.x * kTileW;
  const int nrows = imin(kTileW, N1 - row0);
  float2* Xt = X + (long long)blockIdx.y * N + (long long)row0 * N2;
  // time sample n2 goes to tile position pos2_of_n[n2] (identity unless the plan is prime-factor)
  for (int c = tb; c < nrows; c += nb) {
#pragma unroll 8
    for (int e = tc; e < N2; e += kTW) {
      int pe = e;
      if (pl.gt) pe = __ldg(&pl.fpos2[e]);
      else if constexpr (S::kPfa) pe = __ldg(&pl.pos2_of_n[e]);
      tile[c * P + pe] = Xt[c * N2 + e];
    }
  }
  __syncthreads();
  fwd_stages_smem<S, 0, NS - 1, 1, P>(tile, nrows, pl.s2);
  for (int c = tb; c < nrows; c += nb) {
#pragma unroll 8
    for (int e = tc; e < N2; e += kTW) Xt[c * N2 + e] = tile[c * P + e];
  }
}

// --------------------------------------------------------------------------- registry
// Schedules exactly as fft_plan.h::make_subplan emits them (odd primes descending, then
// powers of two as 16/8/4/2): checked against the runtime plan before use.
using S128 = Sub<128, 8, 16>;
using S256 = Sub<256, 16, 16>;
using S512 = Sub<512, 8, 8, 8>;
using S320 = Sub<320, 5, 8, 8>;
using S165 = Sub<165, 15, 11>;
using S186 = Sub<186, 31, 6>;
using S220 = Sub<220, 11, 20>;
using S279 = Sub<279, 31, 9>;
using S372 = Sub<372, 31, 3, 4>;
using S372b = Sub<372, 31, 12>;          // radix 12 = 4 x 3 in registers: one shared-memory pass fewer
using S440 = Sub<440, 11, 5, 8>;
using S200 = Sub<200, 10, 20>;
using S250 = Sub<250, 10, 25>;
using S248 = Sub<248, 31, 8>;
using S496 = Sub<496, 31, 16>;
using S341 = Sub<341, 31, 11>;           // coprime split of 163680 = 341 x 480
using S480 = Sub<480, 15, 32>;
using S90 = Sub<90, 9, 10>;              // coprime split of 30690 = 341 x 90

template <class S> inline bool schedule_matches(const SubPlan& sp) {
  if (sp.F != S::F || sp.ns != S::NS || (sp.pfa != 0) != S::kPfa) return false;
  for (int j = 0; j < S::NS; ++j)
    if (sp.radix[j] != S::radix(j) || sp.m[j] != S::stride(j)) return false;
  return true;
}

}  // namespace acq
