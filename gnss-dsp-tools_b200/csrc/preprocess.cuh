// Capture front end on the GPU: what every acquire-*.py does between
// reading the file and calling search() — acquire-gps-l1.py:80-96:
//   x = io.get_samples_complex(fp, n)            int8 I/Q -> complex64        (gnsstools/io.py:3-12)
//   nco.mix(x, -coffset/fs, 0)                   carrier wipe-off             (gnsstools/nco.py:30-41)
//   x = scipy.signal.filtfilt(h, [1], x)         161-tap FIR, forward-backward, odd-extension padding
//   xr/xi = np.interp(t/fsr, arange(len(x)), x)  linear-interpolation resample
// Arithmetic follows the reference operation by operation (float32 odd extension, float64
// FIR sums oldest tap first, np.interp's slope form, no FMA contraction): the complex128
// result agrees with scipy/numpy to ~1e-15 relative (summation order inside lfilter) and is
// identical after rounding to the complex64 the search consumes (tests/test_preprocess.py).
#pragma once
#include "kernels.cuh"

namespace acq {

// Mixed sample i of the raw recording as nco.mix leaves it: complex64(int8 pair) times the
// complex128 table entry, rounded to complex64.
__device__ __forceinline__ float2 mixed_sample(const signed char* __restrict__ iq, long long i,
                                               unsigned long long dp0, unsigned long long df,
                                               const double2* __restrict__ tab) {
  const long long dp = (long long)(dp0 + (unsigned long long)i * df);
  const double2 w = tab[(int)((dp >> 50) & (kNcoSize - 1))];
  const double sr = (double)iq[2 * i], si = (double)iq[2 * i + 1];
  const double re = __dadd_rn(__dmul_rn(sr, w.x), -__dmul_rn(si, w.y));
  const double im = __dadd_rn(__dmul_rn(sr, w.y), __dmul_rn(si, w.x));
  return make_float2((float)re, (float)im);
}

// ext = odd_ext(x, edge) in complex64 (scipy.signal._arraytools.odd_ext as filtfilt calls it):
// [2*x[0] - x[edge..1], x, 2*x[n-1] - x[n-2..n-1-edge]].
__global__ void __launch_bounds__(kThreads)
k_pre_extend(const signed char* __restrict__ iq, long long n, int edge, unsigned long long dp0,
             unsigned long long df, const double2* __restrict__ tab, float2* __restrict__ ext) {
  const long long L = n + 2ll * edge;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < L; m += stride) {
    float2 v;
    if (m < edge) {
      const float2 e = mixed_sample(iq, 0, dp0, df, tab), s = mixed_sample(iq, edge - m, dp0, df, tab);
      v = make_float2(__fsub_rn(2.f * e.x, s.x), __fsub_rn(2.f * e.y, s.y));
    } else if (m < edge + n) {
      v = mixed_sample(iq, m - edge, dp0, df, tab);
    } else {
      const long long r = m - edge - n;
      const float2 e = mixed_sample(iq, n - 1, dp0, df, tab), s = mixed_sample(iq, n - 2 - r, dp0, df, tab);
      v = make_float2(__fsub_rn(2.f * e.x, s.x), __fsub_rn(2.f * e.y, s.y));
    }
    ext[m] = v;
  }
}

// One FIR pass in float64, summed oldest tap first like lfilter's transposed direct form:
// DIR=+1 (forward):  y[m] = sum_k b[k] * in[m - k]      for m in [lo, hi)
// DIR=-1 (backward): y[m] = sum_k b[k] * in[m + k]
// Only indices whose taps stay inside the array are evaluated (the caller's [lo, hi)), which
// is all the sliced filtfilt output depends on, so lfilter's initial conditions never enter.
template <int DIR, class TIN>
__global__ void __launch_bounds__(kThreads)
k_pre_fir(const TIN* __restrict__ in, const double* __restrict__ b, int ntaps, long long lo, long long hi,
          double2* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long m = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; m < hi; m += stride) {
    double ar = 0.0, ai = 0.0;
    for (int k = ntaps - 1; k >= 0; --k) {
      const TIN s = in[m - (long long)DIR * k];
      const double bk = b[k];
      const double pr = __dmul_rn(bk, (double)s.x), pi = __dmul_rn(bk, (double)s.y);
      if (k == ntaps - 1) { ar = pr; ai = pi; }
      else { ar = __dadd_rn(ar, pr); ai = __dadd_rn(ai, pi); }
    }
    out[m] = make_double2(ar, ai);
  }
}

// np.interp(step * t, arange(n), z) for t < n_out, real and imaginary parts alike
// (numpy arr_interp: slope * (x - xp[j]) + fp[j], right edge clamps to fp[n-1]); z = zf[edge + .].
__global__ void __launch_bounds__(kThreads)
k_pre_interp(const double2* __restrict__ zf, long long n, int edge, double step, long long n_out,
             float2* __restrict__ out_c64, double2* __restrict__ out_c128) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const double2* z = zf + edge;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_out; t += stride) {
    const double x = __dmul_rn(step, (double)t);
    double re, im;
    if (x >= (double)(n - 1)) { re = z[n - 1].x; im = z[n - 1].y; }
    else {
      const long long j = (long long)floor(x);
      const double dx = __dadd_rn(x, -(double)j);
      const double2 a = z[j], c = z[j + 1];
      re = __dadd_rn(__dmul_rn(__dadd_rn(c.x, -a.x), dx), a.x);
      im = __dadd_rn(__dmul_rn(__dadd_rn(c.y, -a.y), dx), a.y);
    }
    out_c64[t] = make_float2((float)re, (float)im);
    if (out_c128) out_c128[t] = make_double2(re, im);
  }
}

}  // namespace acq
