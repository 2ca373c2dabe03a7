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

}  // namespace acq
