// Mixed-radix complex FFT building blocks that run on a shared-memory tile.
//
// Data model. A *tile* holds `ncols` independent transforms of length F laid out
// batch-fastest: element e of column c lives at tile[e*WP + c]. Butterflies are
// assigned to threads as (column fastest, butterfly next), so a half-warp always touches
// 16 consecutive float2 — bank-conflict-free for every radix and stride, which is what
// lets one routine serve the non-power-of-two lengths GNSS sample rates produce
// (30690 = 2*3^2*5*11*31, 163680 = 2^5*3*5*11*31, 81920 = 2^14*5, ...).
//
// Transform pair. Forward is decimation-in-frequency, in place, natural order in ->
// mixed-radix digit-reversed ("position") order out. Inverse is the exact adjoint
// (stages reversed, conjugate twiddles first, conjugate butterflies), position order in
// -> natural order out, unnormalised. The acquisition path multiplies spectra element by
// element, so the frequency domain never needs natural order and no reordering pass exists.
#pragma once
#include "cuda_compat.h"
#include <type_traits>
#include <utility>

namespace acq {

constexpr int kMaxStages = 10;

// One in-shared-memory transform length and its radix schedule (host-built, see fft_plan.h).
struct SubPlan {
  int F;                    // transform length
  int ns;                   // number of stages
  int radix[kMaxStages];    // radix of stage j (forward order)
  int m[kMaxStages];        // element stride of stage j = product of the later radices
  const float2* tw;         // device table: tw[k] = exp(-2*pi*i*k/F), k < F, followed by the
                            // per-stage tables below
  int tws_off[kMaxStages];  // stage j twiddles, butterfly-major: tw[tws_off[j] + i*(R-1) + (q-1)]
                            //   = exp(-2*pi*i * q*i / (R*m)),  i < m, 1 <= q < R   (0 if m == 1)
  int tws0_t_off;           // stage 0 again, q-major: tw[tws0_t_off + (q-1)*m + i]
  int pfa;                  // 1: prime-factor transform, the stage twiddles are not applied (fft_plan.h)
};

// ---------------------------------------------------------------- complex helpers
__host__ __device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
// Packed FP32x2 arithmetic (sm_100a FFMA2 / FADD2 / FMUL2): one issue slot per complex
// add or real-times-complex FMA instead of two. Same IEEE results as the scalar forms.
#if defined(__CUDA_ARCH__) && !defined(GNSSACQ_NO_F32X2)
__device__ __forceinline__ unsigned long long c2u(float2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 u2c(unsigned long long a) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(a));
  return r;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(c2u(a)), "l"(c2u(b)));
  return u2c(r);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(c2u(a)), "l"(c2u(b)));
  return u2c(r);
}
// acc + s * a  (s real)
__device__ __forceinline__ float2 cfma_real(float s, float2 a, float2 acc) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(c2u(make_float2(s, s))), "l"(c2u(a)), "l"(c2u(acc)));
  return u2c(r);
}
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cfma_real(float s, float2 a, float2 acc) {
  return make_float2(fmaf(s, a.x, acc.x), fmaf(s, a.y, acc.y));
}
#endif
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 cswap(float2 a) { return make_float2(a.y, a.x); }
// a * (-i)
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }

// ---------------------------------------------------------------- compile-time trig
__host__ __device__ constexpr double cx_sin(double x) {
  double term = x, sum = x;
  for (int k = 1; k < 24; ++k) { term *= -x * x / ((2.0 * k) * (2.0 * k + 1.0)); sum += term; }
  return sum;
}
__host__ __device__ constexpr double cx_cos(double x) {
  double term = 1.0, sum = 1.0;
  for (int k = 1; k < 24; ++k) { term *= -x * x / ((2.0 * k - 1.0) * (2.0 * k)); sum += term; }
  return sum;
}
__host__ __device__ constexpr double cx_angle(int k, int n) {
  // 2*pi*k/n reduced to (-pi, pi]
  k %= n;
  double t = 6.283185307179586476925286766559 * (double)k / (double)n;
  return t > 3.14159265358979323846 ? t - 6.283185307179586476925286766559 : t;
}
template <int P> struct TrigTab { float c[P]; float s[P]; };
template <int P> __host__ __device__ constexpr TrigTab<P> make_trig_tab() {
  TrigTab<P> t{};
  for (int k = 0; k < P; ++k) { t.c[k] = (float)cx_cos(cx_angle(k, P)); t.s[k] = (float)cx_sin(cx_angle(k, P)); }
  return t;
}

template <int P> constexpr TrigTab<P> kTrig = make_trig_tab<P>();

template <int I, int N, class F> __device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// ---------------------------------------------------------------- forward butterflies
// dft<R>(v): v[q] <- sum_p v[p] * exp(-2*pi*i*p*q/R), in place, natural order.
template <int R> struct Dft;

template <> struct Dft<2> {
  static __device__ __forceinline__ void run(float2* v) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b); v[1] = csub(a, b);
  }
};
template <> struct Dft<4> {
  static __device__ __forceinline__ void run(float2* v) {
    float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    float2 t2 = cadd(v[1], v[3]), t3 = cmul_mi(csub(v[1], v[3]));
    v[0] = cadd(t0, t2); v[2] = csub(t0, t2);
    v[1] = cadd(t1, t3); v[3] = csub(t1, t3);
  }
};
template <> struct Dft<8> {
  static __device__ __forceinline__ void run(float2* v) {
    const float h = 0.70710678118654752440f;
    float2 e[4] = {v[0], v[2], v[4], v[6]};
    float2 o[4] = {v[1], v[3], v[5], v[7]};
    Dft<4>::run(e); Dft<4>::run(o);
    // o[k] *= W8^k
    o[1] = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));
    o[2] = cmul_mi(o[2]);
    o[3] = make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = cadd(e[k], o[k]); v[k + 4] = csub(e[k], o[k]); }
  }
};
template <> struct Dft<16> {
  static __device__ __forceinline__ void run(float2* v) {
    const float h = 0.70710678118654752440f;
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
    float2 e[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { e[k] = v[2 * k]; o[k] = v[2 * k + 1]; }
    Dft<8>::run(e); Dft<8>::run(o);
    // o[k] *= W16^k = (cos(k*pi/8), -sin(k*pi/8))
    o[1] = cmul(o[1], make_float2(c1, -s1));
    o[2] = make_float2(h * (o[2].x + o[2].y), h * (o[2].y - o[2].x));
    o[3] = cmul(o[3], make_float2(s1, -c1));
    o[4] = cmul_mi(o[4]);
    o[5] = cmul(o[5], make_float2(-s1, -c1));
    o[6] = make_float2(h * (o[6].y - o[6].x), -h * (o[6].x + o[6].y));
    o[7] = cmul(o[7], make_float2(-c1, -s1));
#pragma unroll
    for (int k = 0; k < 8; ++k) { v[k] = cadd(e[k], o[k]); v[k + 8] = csub(e[k], o[k]); }
  }
};

// Odd prime radix: split into symmetric / antisymmetric halves so every output pair
// (k, P-k) shares its real-coefficient sums: (P-1)^2 real FMAs instead of 4*P^2.
template <int P> struct Dft {
  static_assert(P % 2 == 1 && P >= 3, "generic butterfly is for odd radices");
  static __device__ __forceinline__ void run(float2* v) {
    constexpr int H = (P - 1) / 2;
    float2 a[H + 1], b[H + 1];
    const float2 x0 = v[0];
    float2 s0 = x0;
    static_for<1, H + 1>([&](auto J) {
      constexpr int j = decltype(J)::value;
      a[j] = cadd(v[j], v[P - j]);
      b[j] = csub(v[j], v[P - j]);
      s0 = cadd(s0, a[j]);
    });
    v[0] = s0;
    static_for<1, H + 1>([&](auto K) {
      constexpr int k = decltype(K)::value;
      float2 re = x0, im = make_float2(0.f, 0.f);
      static_for<1, H + 1>([&](auto J) {
        constexpr int j = decltype(J)::value;
        constexpr float c = kTrig<P>.c[(j * k) % P];
        constexpr float s = kTrig<P>.s[(j * k) % P];
        re = cfma_real(c, a[j], re);
        im = cfma_real(s, b[j], im);
      });
      v[k] = make_float2(re.x + im.y, re.y - im.x);        // re - i*im
      v[P - k] = make_float2(re.x - im.y, re.y + im.x);    // re + i*im
    });
  }
};

// Composite radices done entirely in registers, so a tile transform needs fewer
// shared-memory passes (372 = 31 * 12 instead of 31 * 3 * 4). Coprime factors use the
// prime-factor (Good-Thomas) index maps — no twiddles at all; 9 and 25 use Cooley-Tukey with
// compile-time twiddles. All indices are compile-time after unrolling (pure register renaming).
__host__ __device__ constexpr int modinv(int a, int m) {
  a %= m;
  for (int x = 1; x < m; ++x)
    if ((a * x) % m == 1) return x;
  return 1;
}

template <int R1, int R2> struct DftPFA {
  static __device__ __forceinline__ void run(float2* v) {
    constexpr int R = R1 * R2;
    constexpr int e1 = R2 * modinv(R2, R1), e2 = R1 * modinv(R1, R2);     // CRT reconstruction
    float2 t[R];
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) {
      float2 u[R1];
#pragma unroll
      for (int n1 = 0; n1 < R1; ++n1) u[n1] = v[(R2 * n1 + R1 * n2) % R];
      Dft<R1>::run(u);
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) t[k1 * R2 + n2] = u[k1];
    }
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
      float2 w[R2];
#pragma unroll
      for (int n2 = 0; n2 < R2; ++n2) w[n2] = t[k1 * R2 + n2];
      Dft<R2>::run(w);
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) v[(k1 * e1 + k2 * e2) % R] = w[k2];
    }
  }
};

template <int R1, int R2> struct DftCT {
  static __device__ __forceinline__ void run(float2* v) {
    constexpr int R = R1 * R2;
    constexpr float kH = 0.70710678118654752440f;
    float2 t[R];
    static_for<0, R2>([&](auto N2) {
      constexpr int n2 = decltype(N2)::value;
      float2 u[R1];
#pragma unroll
      for (int n1 = 0; n1 < R1; ++n1) u[n1] = v[R2 * n1 + n2];
      Dft<R1>::run(u);
      static_for<0, R1>([&](auto K1) {
        constexpr int k1 = decltype(K1)::value;
        constexpr int e = (k1 * n2) % R;                  // u *= W_R^e = exp(-2*pi*i*e/R)
        if constexpr (e == 0) t[k1 * R2 + n2] = u[k1];
        else if constexpr (4 * e == R) t[k1 * R2 + n2] = cmul_mi(u[k1]);
        else if constexpr (2 * e == R) t[k1 * R2 + n2] = make_float2(-u[k1].x, -u[k1].y);
        else if constexpr (4 * e == 3 * R) t[k1 * R2 + n2] = make_float2(-u[k1].y, u[k1].x);
        else if constexpr (8 * e == R) t[k1 * R2 + n2] = make_float2(kH * (u[k1].x + u[k1].y), kH * (u[k1].y - u[k1].x));
        else if constexpr (8 * e == 3 * R) t[k1 * R2 + n2] = make_float2(kH * (u[k1].y - u[k1].x), -kH * (u[k1].x + u[k1].y));
        else if constexpr (8 * e == 5 * R) t[k1 * R2 + n2] = make_float2(-kH * (u[k1].x + u[k1].y), kH * (u[k1].x - u[k1].y));
        else if constexpr (8 * e == 7 * R) t[k1 * R2 + n2] = make_float2(kH * (u[k1].x - u[k1].y), kH * (u[k1].x + u[k1].y));
        else {
          constexpr float wc = kTrig<R>.c[e];
          constexpr float ws = kTrig<R>.s[e];
          t[k1 * R2 + n2] = cmul(u[k1], make_float2(wc, -ws));
        }
      });
    });
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
      float2 w[R2];
#pragma unroll
      for (int n2 = 0; n2 < R2; ++n2) w[n2] = t[k1 * R2 + n2];
      Dft<R2>::run(w);
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = w[k2];
    }
  }
};

// 32 = 4 x 8 Cooley-Tukey with compile-time twiddles: lets 480 = 15 * 32 and 352 = 11 * 32 run as
// two-stage prime-factor tile transforms (one shared-memory pass).
template <> struct Dft<32> : DftCT<4, 8> {};
template <> struct Dft<6> : DftPFA<2, 3> {};
template <> struct Dft<10> : DftPFA<2, 5> {};
template <> struct Dft<12> : DftPFA<4, 3> {};
template <> struct Dft<15> : DftPFA<3, 5> {};
template <> struct Dft<20> : DftPFA<4, 5> {};
template <> struct Dft<22> : DftPFA<2, 11> {};
template <> struct Dft<9> : DftCT<3, 3> {};
template <> struct Dft<25> : DftCT<5, 5> {};

// ---------------------------------------------------------------- one stage over a tile
// Threads are viewed as (tc = tid % TW, tb = tid / TW): tc walks columns, tb walks butterflies.
constexpr int kTW = 16;

template <int R, bool INV>
__device__ __forceinline__ void stage_tile(float2* tile, int WP, int ncols, int F, int m,
                                           const float2* __restrict__ tw) {
  const int tc = threadIdx.x & (kTW - 1);
  const int tb = threadIdx.x / kTW;
  const int nb = blockDim.x / kTW;
  const int nbf = F / R;
  const int twstep = F / (R * m);
  const int estride = m * WP;
  for (int bf = tb; bf < nbf; bf += nb) {
    const int blk = bf / m;
    const int i = bf - blk * m;
    float2* base = tile + (blk * R * m + i) * WP;
    constexpr bool kCacheTw = R <= 8;      // larger radices reload (broadcast, L1-resident)
    float2 w[kCacheTw ? R : 1];
    const float2* twp = tw + 0;
    const int twi = i * twstep;
    if (kCacheTw && m > 1) {
#pragma unroll
      for (int q = 1; q < R; ++q) w[q] = __ldg(&twp[q * twi]);
    }
    for (int c = tc; c < ncols; c += kTW) {
      float2* p = base + c;
      float2 v[R];
#pragma unroll
      for (int q = 0; q < R; ++q) v[q] = p[q * estride];
      if (INV) {
        if (m > 1) {
#pragma unroll
          for (int q = 1; q < R; ++q) v[q] = cmulc(v[q], kCacheTw ? w[q] : __ldg(&twp[q * twi]));
        }
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = cswap(v[q]);
        Dft<R>::run(v);
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = cswap(v[q]);
      } else {
        Dft<R>::run(v);
        if (m > 1) {
#pragma unroll
          for (int q = 1; q < R; ++q) v[q] = cmul(v[q], kCacheTw ? w[q] : __ldg(&twp[q * twi]));
        }
      }
#pragma unroll
      for (int q = 0; q < R; ++q) p[q * estride] = v[q];
    }
  }
}

// Large odd prime radix (31): the symmetric-sum butterfly needs ~4P live registers in one
// thread, which would cap the whole kernel at one CTA per SM. Instead two warps share each
// butterfly: both load all P inputs and form the a/b halves, then the even warp produces
// outputs {0, 1..K0, P-1..P-K0} and the odd warp the rest — no exchange, half the
// accumulators each. A block barrier separates the loads from the in-place stores.
// Output pairs (k, P-k) for k = K0..K1, produced kPrimeBatch pairs at a time: each pair is two
// dependent chains of H packed FMAs, so a batch keeps 2 * kPrimeBatch independent chains in flight.
// batch(k0, nk, re, im) receives integral constants k0, nk and the accumulators of pairs
// k0 .. k0+nk-1: output k = re - i*im, output P-k = re + i*im.
#ifdef GNSSACQ_PRIME_BATCH
constexpr int kPrimeBatch = GNSSACQ_PRIME_BATCH;      // A/B builds of tools/microbench only
#else
constexpr int kPrimeBatch = 3;
#endif
template <int P, int K0, int K1, class Batch>
__device__ __forceinline__ void prime_outputs_batched(const float2 x0, const float2* a, const float2* b, Batch&& batch) {
  constexpr int H = (P - 1) / 2;
  constexpr int NK = K1 - K0 + 1;
  static_for<0, (NK + kPrimeBatch - 1) / kPrimeBatch>([&](auto G) {
    constexpr int k0 = K0 + decltype(G)::value * kPrimeBatch;
    constexpr int nk = (K1 - k0 + 1) < kPrimeBatch ? (K1 - k0 + 1) : kPrimeBatch;
    float2 re[nk], im[nk];
#pragma unroll
    for (int t = 0; t < nk; ++t) { re[t] = x0; im[t] = make_float2(0.f, 0.f); }
    static_for<1, H + 1>([&](auto J) {
      constexpr int j = decltype(J)::value;
      static_for<0, nk>([&](auto T) {
        constexpr int t = decltype(T)::value, k = k0 + t;
        constexpr float c = kTrig<P>.c[(j * k) % P];
        constexpr float sn = kTrig<P>.s[(j * k) % P];
        re[t] = cfma_real(c, a[j], re[t]);
        im[t] = cfma_real(sn, b[j], im[t]);
      });
    });
    batch(std::integral_constant<int, k0>{}, std::integral_constant<int, nk>{}, re, im);
  });
}
template <int P, int K0, int K1, class Emit>
__device__ __forceinline__ void prime_outputs(const float2 x0, const float2* a, const float2* b, Emit&& emit) {
  prime_outputs_batched<P, K0, K1>(x0, a, b, [&](auto K0c, auto NKc, const float2* re, const float2* im) {
    constexpr int k0 = decltype(K0c)::value, nk = decltype(NKc)::value;
    static_for<0, nk>([&](auto T) {
      constexpr int t = decltype(T)::value, k = k0 + t;
      emit(k, make_float2(re[t].x + im[t].y, re[t].y - im[t].x));
      emit(P - k, make_float2(re[t].x - im[t].y, re[t].y + im[t].x));
    });
  });
}

// ---------------------------------------------------------------- radix 31 as two 15-point cyclic convolutions
// The symmetric-sum butterfly above spends (P-1)^2 = 900 real FMAs on 31 points: two 15 x 15 real matrices
// (cos on a[j] = x[j] + x[31-j], sin on b[j] = x[j] - x[31-j]) applied to complex vectors. Both matrices are
// convolutions in disguise (Rader): 3 generates (Z/31)*, and modulo +-1 that group is cyclic of order 15, so
// with class index m (representative jr[m] = +-3^m mod 31 in 1..15)
//     cos(2 pi jr[m] jr[n] / 31) = hc[(m + n) mod 15],         hc[t] = cos(2 pi 3^t / 31)
//     sin(2 pi jr[m] jr[n] / 31) = sg[m] sg[n] hs[(m + n) mod 15], hs[t] = (-1)^t sin(2 pi 3^t / 31),
// sg[m] = sign(3^m mod 31 <= 15) (-1)^m (3^15 = -1 makes the sine sequence anti-periodic; 15 is odd, so the
// alternating sign turns it periodic). Signs are free: b[j] is formed with its operands swapped, and a
// negative output sign exchanges outputs k and 31 - k. A 15-point cyclic convolution is a 3 x 5
// two-dimensional one (CRT on the index); along the 3-point axis Winograd's algorithm needs 4 block products
// instead of 9, the blocks being 5 x 5 matrix-vector products done with FMAs:
//     S = x0 + x1 + x2, u = x0 - x2, v = x1 - x2, w = u + v                       (x_i: 5-vectors, 25 adds)
//     t = Hs S,  m = t + H1' w,  y0 = m + (H0' - H1') u,  y1 = m + (H2' - H1') v    (4 x 25 FMAs)
//     y2 = 3 t - y0 - y1                                                           (10 operations)
// with Hs = (H0 + H1 + H2) / 3 and Hi' = Hi - Hs: 135 packed operations per convolution instead of 225,
// 270 + 5 per butterfly instead of 450 + 15. Block 2 is delivered NEGATED in both convolutions
// (-(y0 - 3 t + y1): no packed negation exists), i.e. outputs of block 2 are the negatives of the DFT outputs —
// the callers take |.| of them.
__host__ __device__ constexpr int r31_crt(int a3, int a5) { return (10 * a3 + 6 * a5) % 15; }
__host__ __device__ constexpr int r31_pow3(int m) { int g = 1; for (int i = 0; i < m; ++i) g = (g * 3) % 31; return g; }
__host__ __device__ constexpr int r31_rep(int m) { const int g = r31_pow3(m); return g <= 15 ? g : 31 - g; }          // jr[m]
__host__ __device__ constexpr int r31_sign(int m) { return ((r31_pow3(m) <= 15) ? 1 : -1) * ((m & 1) ? -1 : 1); }      // sg[m]
__host__ __device__ constexpr int r31_class(int j) { for (int m = 0; m < 15; ++m) if (r31_rep(m) == j) return m; return -1; }
// b[j] must be handed over as sg[class(j)] * (x[j] - x[31-j]): true when the operands have to be swapped
__host__ __device__ constexpr bool r31_flip_b(int j) { return r31_sign(r31_class(j)) < 0; }
// digit (output index of the 31-point transform) of value t of block n3: values come in pairs, pair p = t / 2 is
// class n = crt(n3, p); the even value is re + i*im' (im' = the sine convolution as delivered), the odd one re - i*im'
__host__ __device__ constexpr int r31_digit(int n3, int t) {
  const int n = r31_crt(n3, t / 2), k = r31_rep(n);
  return ((t & 1) == 0) == (r31_sign(n) > 0) ? k : 31 - k;      // caller's convention: output k = re + i*im_k, im_k = sg[n] im'
}
#ifdef GNSSACQ_NO_RADER31
constexpr bool kRader31On = false;            // A/B builds: the (P-1)^2 butterfly everywhere
#else
constexpr bool kRader31On = true;
#endif
// compile-time digit lists handed to the epilogues
template <int N> struct IntArray { int v[N]; int n; };
template <int... Q> __host__ __device__ constexpr IntArray<sizeof...(Q)> seq_array(std::integer_sequence<int, Q...>) {
  return IntArray<sizeof...(Q)>{{Q...}, (int)sizeof...(Q)};
}
template <int N3, int... T> __host__ __device__ constexpr auto r31_block_digits(std::integer_sequence<int, T...>) {
  return std::integer_sequence<int, r31_digit(N3, T)...>{};
}
// pairs k0, k0+1, ...: even value = output k, odd value = output P - k
template <int P, int K0, int... T> __host__ __device__ constexpr auto prime_pair_digits(std::integer_sequence<int, T...>) {
  return std::integer_sequence<int, ((T & 1) ? P - (K0 + T / 2) : K0 + T / 2)...>{};
}
struct Rader31Tab { float hc[4][5], hs[4][5]; };
__host__ __device__ constexpr Rader31Tab make_rader31_tab() {
  Rader31Tab r{};
  for (int pass = 0; pass < 2; ++pass)
    for (int t5 = 0; t5 < 5; ++t5) {
      double h[3] = {0.0, 0.0, 0.0};
      for (int t3 = 0; t3 < 3; ++t3) {
        const int t = r31_crt(t3, t5), g = r31_pow3(t);
        h[t3] = pass ? ((t & 1) ? -1.0 : 1.0) * cx_sin(cx_angle(g, 31)) : cx_cos(cx_angle(g, 31));
      }
      const double s = (h[0] + h[1] + h[2]) / 3.0;
      float (*o)[5] = pass ? r.hs : r.hc;
      o[0][t5] = (float)s;
      o[1][t5] = (float)(h[1] - s);
      o[2][t5] = (float)(h[0] - h[1]);
      o[3][t5] = (float)(h[2] - h[1]);
    }
  return r;
}
template <int P> constexpr Rader31Tab kRader31 = make_rader31_tab();

// y[n5] = init[n5] + sum_m5 M[K][(m5 + n5) mod 5] x[m5]: five interleaved chains of five packed FMAs
template <bool SIN, int K>
__device__ __forceinline__ void r31_block(const float2* x, const float2* init, float2* y) {
#pragma unroll
  for (int n5 = 0; n5 < 5; ++n5) y[n5] = init[n5];
  static_for<0, 5>([&](auto M5) {
    static_for<0, 5>([&](auto N5) {
      constexpr int m5 = decltype(M5)::value, n5 = decltype(N5)::value;
      constexpr float c = SIN ? kRader31<31>.hs[K][(m5 + n5) % 5] : kRader31<31>.hc[K][(m5 + n5) % 5];
      y[n5] = cfma_real(c, x[m5], y[n5]);
    });
  });
}
// One convolution. in(integral_constant m) = input of class m; blk(integral_constant n3, y) receives block n3
// (classes crt(n3, 0..4)), block 2 negated; sums(S) sees the five column sums (their total is sum_m in(m)).
template <bool SIN, class In, class Sums, class Blk>
__device__ __forceinline__ void r31_conv(const float2 init, In&& in, Sums&& sums, Blk&& blk) {
  float2 S[5], u[5], v[5], w[5];
  static_for<0, 5>([&](auto M5) {
    constexpr int m5 = decltype(M5)::value;
    const float2 x0 = in(std::integral_constant<int, r31_crt(0, m5)>{});
    const float2 x1 = in(std::integral_constant<int, r31_crt(1, m5)>{});
    const float2 x2 = in(std::integral_constant<int, r31_crt(2, m5)>{});
    S[m5] = cadd(cadd(x0, x1), x2);
    u[m5] = csub(x0, x2);
    v[m5] = csub(x1, x2);
    w[m5] = cadd(u[m5], v[m5]);
  });
  sums(S);
  float2 i5[5], t[5], m[5], y[5], d[5];
#pragma unroll
  for (int n5 = 0; n5 < 5; ++n5) i5[n5] = init;
  r31_block<SIN, 0>(S, i5, t);
  r31_block<SIN, 1>(w, t, m);
  r31_block<SIN, 2>(u, m, y);
  blk(std::integral_constant<int, 0>{}, y);
#pragma unroll
  for (int n5 = 0; n5 < 5; ++n5) d[n5] = cfma_real(-3.f, t[n5], y[n5]);      // y0 - 3 t
  r31_block<SIN, 3>(v, m, y);
  blk(std::integral_constant<int, 1>{}, y);
#pragma unroll
  for (int n5 = 0; n5 < 5; ++n5) y[n5] = cadd(d[n5], y[n5]);                 // -(y2)
  blk(std::integral_constant<int, 2>{}, y);
}
// All 31 outputs of the butterfly whose inputs are x0, a[j] = x[j] + x[31-j] and
// bs[j] = (r31_flip_b(j) ? x[31-j] - x[j] : x[j] - x[31-j]), j = 1..15.
// dc(s0): output 0 = x0 + sum a[j]. batch(integral_constant n3, re, im): for p = 0..4 the pair
// (re[p] + i*im[p], re[p] - i*im[p]) in the (re.x - im.y, re.y + im.x) sense are the outputs with digits
// r31_digit(n3, 2p), r31_digit(n3, 2p + 1) — negated for n3 == 2.
template <class DC, class Batch>
__device__ __forceinline__ void rader31_outputs(const float2 x0, const float2* a, const float2* bs, DC&& dc, Batch&& batch) {
  float2 yc[3][5];
  r31_conv<false>(x0, [&](auto M) { return a[r31_rep(decltype(M)::value)]; },
                  [&](const float2* S) { dc(cadd(cadd(cadd(x0, S[0]), cadd(S[1], S[2])), cadd(S[3], S[4]))); },
                  [&](auto N3, const float2* y) {
#pragma unroll
                    for (int p = 0; p < 5; ++p) yc[decltype(N3)::value][p] = y[p];
                  });
  r31_conv<true>(make_float2(0.f, 0.f), [&](auto M) { return bs[r31_rep(decltype(M)::value)]; }, [](const float2*) {},
                 [&](auto N3, const float2* y) { batch(N3, yc[decltype(N3)::value], y); });
}

// The whole forward 31-point transform of one thread's registers through the two convolutions (outputs in natural
// order, block 2's negation undone by negated additions).
struct Dft31Rader {
  static __device__ __forceinline__ void run(float2* v) {
    float2 a[16], bs[16];
    const float2 x0 = v[0];
    static_for<1, 16>([&](auto J) {
      constexpr int j = decltype(J)::value;
      const float2 u = v[j], w = v[31 - j];
      a[j] = cadd(u, w);
      bs[j] = r31_flip_b(j) ? csub(w, u) : csub(u, w);
    });
    rader31_outputs(x0, a, bs, [&](const float2 dc) { v[0] = dc; },
                    [&](auto N3, const float2* re, const float2* im) {
                      constexpr int n3 = decltype(N3)::value;
                      static_for<0, 5>([&](auto Pp) {
                        constexpr int p = decltype(Pp)::value, n = r31_crt(n3, p), k = r31_rep(n);
                        constexpr bool pos = r31_sign(n) > 0;
                        float2 u, w;                             // re - i*im', re + i*im' (forward sign), im' as delivered
                        if constexpr (n3 == 2) {
                          u = make_float2(-re[p].x - im[p].y, im[p].x - re[p].y);
                          w = make_float2(im[p].y - re[p].x, -re[p].y - im[p].x);
                        } else {
                          u = make_float2(re[p].x + im[p].y, re[p].y - im[p].x);
                          w = make_float2(re[p].x - im[p].y, re[p].y + im[p].x);
                        }
                        v[pos ? k : 31 - k] = u;
                        v[pos ? 31 - k : k] = w;
                      });
                    });
  }
};

// NOTW: prime-factor transform, no stage twiddles.
template <int P, bool INV, int ES = 0, int CS = 1, bool NOTW = false>
__device__ __forceinline__ void stage_tile_split(float2* tile, int ncols, int F, int m,
                                                 const float2* __restrict__ tw, int WPrt = 0) {
  const int WP = ES ? ES : WPrt;             // element stride: compile-time when ES != 0
  constexpr int H = (P - 1) / 2;
  constexpr int KA = (H + 1) / 2;            // even warp: k = 1..KA (+ output 0); odd warp: KA+1..H
  const int tc = threadIdx.x & (kTW - 1);
  const int warp = threadIdx.x >> 5;
  const int role = warp & 1;
  const int slot = (warp >> 1) * 2 + ((threadIdx.x >> 4) & 1);
  const int nslots = blockDim.x >> 5;        // butterflies in flight per round
  const int nbf = F / P;
  const int twstep = F / (P * m);
  const int estride = m * WP;
  const int rounds = (nbf + nslots - 1) / nslots;
  for (int c0 = 0; c0 < ncols; c0 += kTW)    // uniform trip counts: every thread meets every barrier
  for (int rd = 0; rd < rounds; ++rd) {
    const int bf = rd * nslots + slot;
    const bool active = bf < nbf && c0 + tc < ncols;
    const int blk = bf / m;
    const int i = bf - blk * m;
    float2* p = tile + (blk * P * m + i) * WP + (c0 + tc) * CS;
    const int twi = i * twstep;
    float2 a[H + 1], b[H + 1];
    float2 x0 = make_float2(0.f, 0.f);
    auto in = [&](int q) -> float2 {
      float2 v = p[q * estride];
      if (INV) {
        if (!NOTW && m > 1 && q > 0) v = cmulc(v, __ldg(&tw[q * twi]));
        v = cswap(v);
      }
      return v;
    };
    if (active) {
      x0 = in(0);
      static_for<1, H + 1>([&](auto J) {
        constexpr int j = decltype(J)::value;
        const float2 u = in(j), w = in(P - j);
        a[j] = cadd(u, w);
        b[j] = csub(u, w);
      });
    }
    __syncthreads();                         // every input read before any in-place store
    if (active) {
      auto emit = [&](int q, float2 v) {
        if (INV) v = cswap(v);
        else if (!NOTW && m > 1 && q > 0) v = cmul(v, __ldg(&tw[q * twi]));
        p[q * estride] = v;
      };
      if (role == 0) {
        float2 s0 = x0;
        static_for<1, H + 1>([&](auto J) { s0 = cadd(s0, a[decltype(J)::value]); });
        emit(0, s0);
        prime_outputs<P, 1, KA>(x0, a, b, emit);
      } else {
        prime_outputs<P, KA + 1, H>(x0, a, b, emit);
      }
    }
  }
}

// Radix classes: kernels are instantiated per class so that power-of-two plans are not
// register-allocated for the 31-point butterfly. 0: {2,4,8,16}; 1: + {3,5,6,9,10,12,15,20,25};
// 2: + {7,11,13,22,31,32}.
constexpr int kNumRadixClasses = 3;
__host__ __device__ constexpr int radix_class_of(int R) {
  return (R == 2 || R == 4 || R == 8 || R == 16) ? 0
       : ((R == 3 || R == 5 || R == 6 || R == 9 || R == 10 || R == 12 || R == 15 || R == 20 || R == 25) ? 1 : 2);   // 32 -> 2
}
// prime factors the planner accepts
__host__ __device__ constexpr bool radix_supported(int R) {
  return R == 2 || R == 3 || R == 4 || R == 5 || R == 7 || R == 8 || R == 11 || R == 13 || R == 16 || R == 31;
}

template <int RC, bool INV>
__device__ __forceinline__ void stage_dispatch(int R, float2* tile, int WP, int ncols, int F, int m,
                                               const float2* __restrict__ tw) {
  switch (R) {
    case 2: stage_tile<2, INV>(tile, WP, ncols, F, m, tw); break;
    case 4: stage_tile<4, INV>(tile, WP, ncols, F, m, tw); break;
    case 8: stage_tile<8, INV>(tile, WP, ncols, F, m, tw); break;
    case 16: stage_tile<16, INV>(tile, WP, ncols, F, m, tw); break;
    default:
      if constexpr (RC >= 1) {
        switch (R) {
          case 3: stage_tile<3, INV>(tile, WP, ncols, F, m, tw); break;
          case 5: stage_tile<5, INV>(tile, WP, ncols, F, m, tw); break;
          case 6: stage_tile<6, INV>(tile, WP, ncols, F, m, tw); break;
          case 9: stage_tile<9, INV>(tile, WP, ncols, F, m, tw); break;
          case 10: stage_tile<10, INV>(tile, WP, ncols, F, m, tw); break;
          case 12: stage_tile<12, INV>(tile, WP, ncols, F, m, tw); break;
          case 15: stage_tile<15, INV>(tile, WP, ncols, F, m, tw); break;
          case 20: stage_tile<20, INV>(tile, WP, ncols, F, m, tw); break;
          case 25: stage_tile<25, INV>(tile, WP, ncols, F, m, tw); break;
          default:
            if constexpr (RC >= 2) {
              switch (R) {
                case 7: stage_tile<7, INV>(tile, WP, ncols, F, m, tw); break;
                case 11: stage_tile<11, INV>(tile, WP, ncols, F, m, tw); break;
                case 13: stage_tile<13, INV>(tile, WP, ncols, F, m, tw); break;
                case 22: stage_tile<22, INV>(tile, WP, ncols, F, m, tw); break;
                case 31: stage_tile_split<31, INV>(tile, ncols, F, m, tw, WP); break;
                case 32: stage_tile<32, INV>(tile, WP, ncols, F, m, tw); break;
                default: break;   // the host planner never emits other radices
              }
            }
            break;
        }
      }
      break;
  }
}

// Transform every column of the tile. The caller must have synchronised after filling the
// tile; on return all threads have passed a barrier after the last stage.
template <int RC, bool INV>
__device__ __forceinline__ void subfft_tile(float2* tile, int WP, int ncols, const SubPlan& sp) {
  if (!INV) {
    for (int j = 0; j < sp.ns; ++j) {
      stage_dispatch<RC, false>(sp.radix[j], tile, WP, ncols, sp.F, sp.m[j], sp.tw);
      __syncthreads();
    }
  } else {
    for (int j = sp.ns - 1; j >= 0; --j) {
      stage_dispatch<RC, true>(sp.radix[j], tile, WP, ncols, sp.F, sp.m[j], sp.tw);
      __syncthreads();
    }
  }
}

}  // namespace acq
