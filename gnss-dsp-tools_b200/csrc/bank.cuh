// Replica builder and time-domain correlator bank.
//
// Both evaluate the reference's code resampler on the device,
//   idx = floor((chips % L) + frac + incr * i) mod L            gnsstools/gps/ca.py:106-112
// with the reference's float64 rounding sequence: base = (chips % L) + frac is formed on the
// host (two float64 operations, as numpy does), the device adds the separately rounded
// product incr * i and floors — so every sample picks the chip numpy picks.
#pragma once
#include "kernels.cuh"

namespace acq {

__device__ __forceinline__ int resample_index(double base, double incr, int i, int L) {
  const double v = __dadd_rn(base, __dmul_rn(incr, (double)i));
  long long k = __double2ll_rd(v);
  k %= (long long)L;
  if (k < 0) k += L;                                   // np.mod: result has the divisor's sign
  return (int)k;
}

// ---------------------------------------------------------------- replica builder
// rep[r][i] = (1 - 2 c_r[idx(i)]) * boc(i) for i < n, 0 for n <= i < N   (float32, +-1 / 0):
// acquire-gps-l1.py:22-24 (plain), acquire-gps-l1cd.py:22-26 (x BOC(1,1)),
// acquire-gps-l5i.py:22-24 (zero half). BOC(1,1) = c[floor(2*((chips % 2) + frac + incr*i)) mod 2],
// c = [-1, 1] (gnsstools/nco.py:12-19); base2 = (chips % 2) + frac comes from the host.
// grid = (ceil(N / 256), R)
__global__ void __launch_bounds__(kThreads)
k_build_replicas(const signed char* __restrict__ chips01, int L, int n, int N, double base, double incr,
                 int boc, double base2, float* __restrict__ rep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int r = blockIdx.y;
  float v = 0.f;
  if (i < n) {
    const int idx = resample_index(base, incr, i, L);
    v = 1.0f - 2.0f * (float)chips01[(long long)r * L + idx];
    if (boc) {
      const double ph = __dadd_rn(base2, __dmul_rn(incr, (double)i)) * 2.0;      // exact scaling
      const long long k = __double2ll_rd(ph);
      if ((k & 1ll) == 0) v = -v;                      // c = [-1, 1]
    }
  }
  rep[(long long)r * N + i] = v;
}

// ---------------------------------------------------------------- correlator bank
// The serial long-code acquisitions (acquire-gps-l2cl.py:18-33, acquire-glonass-l1-p.py:14-32,
// acquire-glonass-l2-p.py) test H code-phase hypotheses by direct correlation:
//   P[h][b] = sum_i x[b*stride + i] * w[i] * (1 - 2 c[idx_{h,b}(i)]),  w = nco(f, 0, n)
//   q[h] = sum_b |P[h][b]|
// One wipe-off pass forms xw = x * w once (shared by every hypothesis); then one CTA per
// (hypothesis, block) streams xw and gathers chips. Partial sums are float32 per thread over
// n / 256 terms, folded in float64.
// grid = (ceil(n / 256), B)
__global__ void __launch_bounds__(kThreads)
k_bank_wipeoff(const float2* __restrict__ x, const float2* __restrict__ nco_tab, double f, int n, int stride,
               float2* __restrict__ xw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = blockIdx.y;
  xw[(long long)b * n + i] = cmul(__ldg(&x[(long long)b * stride + i]), nco_sample(nco_tab, f, i));
}

// grid = (H, B); base[h*B + b] = (chips_{h,b} % L) + frac
__global__ void __launch_bounds__(kThreads)
k_corr_bank(const float2* __restrict__ xw, const signed char* __restrict__ chips01, int L, int n, int B,
            const double* __restrict__ base, double incr, double2* __restrict__ out) {
  const int h = blockIdx.x, b = blockIdx.y;
  const double bs = base[(long long)h * B + b];
  const float2* xb = xw + (long long)b * n;
  float sr = 0.f, si = 0.f;
  for (int i = threadIdx.x; i < n; i += kThreads) {
    const int idx = resample_index(bs, incr, i, L);
    const float2 v = __ldg(&xb[i]);
    const float s = 1.0f - 2.0f * (float)__ldg(&chips01[idx]);
    sr = fmaf(s, v.x, sr);
    si = fmaf(s, v.y, si);
  }
  double dr = (double)sr, di = (double)si;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dr += __shfl_xor_sync(0xffffffffu, dr, o);
    di += __shfl_xor_sync(0xffffffffu, di, o);
  }
  __shared__ double s_r[kThreads / 32], s_i[kThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_r[warp] = dr; s_i[warp] = di; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ar = 0.0, ai = 0.0;
    for (int k = 0; k < kThreads / 32; ++k) { ar += s_r[k]; ai += s_i[k]; }
    out[(long long)h * B + b] = make_double2(ar, ai);
  }
}

// ---------------------------------------------------------------- tracking correlators (E/P/L)
// <sig>.correlate(x, prn, chips, frac, incr, c[, boc11]) of the tracking scripts — plain
// (gnsstools/gps/ca.py:120-128), two-level sub-chip pattern (BOC(1,1) data channels, L2C RZ slots),
// CBOC (gnsstools/galileo/e1b.py:45-58), TMBOC (gnsstools/gps/l1cp.py:210-228) — for H hypotheses
// (channel x early/prompt/late tap) at once. The reference advances its code phase by a float64
// recurrence, cp <- (cp + incr) mod L, whose rounding is part of which chip a sample sees; one
// lane per phase accumulator therefore runs the recurrence itself, a chunk of samples at a time,
// into shared index arrays, and the whole CTA then forms the products of that chunk and sums them
// in float64 (the reference sums in sample order; the sum agrees to ~1e-15 relative, every chip
// index is the reference's). Data-parallel across hypotheses (one CTA each), sequential in time.
constexpr int kEplChunk = 2048;

// Python's float modulo for a positive modulus: fmod, moved into [0, m)
__device__ __forceinline__ double py_mod(double a, double m) {
  double r = fmod(a, m);
  if (r != 0.0 && r < 0.0) r += m;
  return r;
}
// (t) mod m for 0 <= t, by exact subtractions (t - m is exact for t >= m)
__device__ __forceinline__ double wrap_mod(double t, double m) {
  while (t >= m) t -= m;
  return t;
}

// mode 0: plain; 1: sub[int(bp)]; 2: a1*sub[int(bp)] + a6*sub[int(bp6)]; 3: pattern[int(cp) % 33] ? sub[int(bp6)] : sub[int(bp)]
// params: {sub0, sub1, a1, a6, pattern[0..32]}. grid = H, block = kThreads.
__global__ void __launch_bounds__(kThreads)
k_correlate_epl(const float2* __restrict__ x, int n, const signed char* __restrict__ chips01, int L, int mode,
                const double* __restrict__ params, const int* __restrict__ xsel, const int* __restrict__ csel,
                const double* __restrict__ start, const double* __restrict__ incr, double2* __restrict__ out) {
  __shared__ int s_cp[kEplChunk];
  __shared__ unsigned char s_bp[kEplChunk], s_bp6[kEplChunk];
  __shared__ double s_re[kThreads / 32], s_im[kThreads / 32];
  const int h = blockIdx.x, tid = threadIdx.x;
  const float2* xb = x + (long long)xsel[h] * n;
  const signed char* code = chips01 + (long long)csel[h] * L;
  const double st = start[h], inc = incr[h], Ld = (double)L;
  // each accumulator lives in one lane (of different warps, so they run side by side)
  double ph = 0.0, step = 0.0, m = 2.0;
  if (tid == 0) { ph = py_mod(st, Ld); step = inc; m = Ld; }                                  // cp = (chips+frac) % L
  if (tid == 32) { ph = py_mod(2.0 * st, 2.0); step = 2.0 * inc; }                             // bp = (2*(chips+frac)) % 2
  if (tid == 64) { ph = py_mod(__dmul_rn(12.0, st), 2.0); step = __dmul_rn(12.0, inc); }       // bp6 = (12*(chips+frac)) % 2
  const bool seq = tid == 0 || (mode >= 1 && tid == 32) || (mode >= 2 && tid == 64);
  const double sub0 = mode >= 1 ? params[0] : 1.0, sub1 = mode >= 1 ? params[1] : 1.0;
  const double a1 = mode == 2 ? params[2] : 0.0, a6 = mode == 2 ? params[3] : 0.0;
  double re = 0.0, im = 0.0;
  for (int i0 = 0; i0 < n; i0 += kEplChunk) {
    const int cnt = imin(kEplChunk, n - i0);
    if (seq) {
      for (int i = 0; i < cnt; ++i) {
        const int k = (int)ph;                                    // int(): truncation of a non-negative value
        if (tid == 0) s_cp[i] = k; else if (tid == 32) s_bp[i] = (unsigned char)k; else s_bp6[i] = (unsigned char)k;
        ph = wrap_mod(__dadd_rn(ph, step), m);                    // (ph + step) % m
      }
    }
    __syncthreads();
    for (int i = tid; i < cnt; i += kThreads) {
      const int k = s_cp[i];
      double coef = 1.0 - 2.0 * (double)code[k];
      if (mode == 1) coef *= s_bp[i] ? sub1 : sub0;
      else if (mode == 2) coef *= __dadd_rn(__dmul_rn(a1, s_bp[i] ? sub1 : sub0), __dmul_rn(a6, s_bp6[i] ? sub1 : sub0));
      else if (mode == 3) coef *= (params[4 + k % 33] != 0.0) ? (s_bp6[i] ? sub1 : sub0) : (s_bp[i] ? sub1 : sub0);
      const float2 v = xb[i0 + i];
      re += (double)v.x * coef;
      im += (double)v.y * coef;
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
  if ((tid & 31) == 0) { s_re[tid >> 5] = re; s_im[tid >> 5] = im; }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) { a += s_re[w]; b += s_im[w]; }
    out[h] = make_double2(a, b);
  }
}

}  // namespace acq
