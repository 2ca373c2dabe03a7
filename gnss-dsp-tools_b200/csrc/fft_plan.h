// Host-side FFT planning: factor the transform length, choose the two-level split
// N = N1 * N2, and build the twiddle tables (computed in double, stored as float2).
#pragma once
#include "fft_core.cuh"
#include <algorithm>
#include <functional>
#include <string>
#include <vector>

namespace acq {

constexpr int kMaxSub = 1024;        // longest transform done inside one shared-memory tile
constexpr int kMaxPrime = 31;        // largest radix the kernels instantiate
constexpr int kMidMax = 8192;        // N <= kMidMax: whole transform resident in one CTA

struct HostSubPlan {
  int F = 0;
  std::vector<int> radix, m;
  std::vector<int> tws_off;          // offset of each stage's butterfly-major twiddle table
  int tws0_t_off = 0;                // offset of stage 0's q-major copy
  std::vector<int> pos_of_freq;      // position of frequency k after the forward transform
  std::vector<int> freq_of_pos;
  // Prime-factor (Good-Thomas) form: when the radices are pairwise coprime the transform is a
  // plain multi-dimensional DFT over the digits of the tile position — no stage twiddles — at the
  // price of index maps on both sides: position p = sum_j d_j m_j holds time sample
  // n = sum_j d_j (F/R_j) mod F before, and frequency k with k = d_j (mod R_j) after.
  bool pfa = false;
  std::vector<int> n_of_pos, pos_of_n;   // identity unless pfa
};

inline bool coprime_schedule(const std::vector<int>& r) {
  if (r.size() < 2) return false;
  auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
  for (size_t i = 0; i < r.size(); ++i)
    for (size_t j = i + 1; j < r.size(); ++j)
      if (gcd(r[i], r[j]) != 1) return false;
  return true;
}

// Switch a sub-plan to the prime-factor index maps (the radix schedule and strides stay).
inline void set_pfa(HostSubPlan& sp) {
  const int F = sp.F;
  sp.pfa = true;
  for (int p = 0; p < F; ++p) {
    long long n = 0, k = 0;
    for (size_t j = 0; j < sp.radix.size(); ++j) {
      const int R = sp.radix[j], d = (p / sp.m[j]) % R, Nj = F / R;
      n += (long long)d * Nj;
      k += (long long)d * Nj * modinv(Nj, R);          // CRT: k = d_j (mod R_j)
    }
    sp.n_of_pos[p] = (int)(n % F);
    sp.pos_of_n[(int)(n % F)] = p;
    sp.freq_of_pos[p] = (int)(k % F);
    sp.pos_of_freq[(int)(k % F)] = p;
  }
}

struct HostPlan {
  int N = 0, N1 = 0, N2 = 0;
  bool large = false;                // two kernels through an L2-resident scratch
  int rclass = 0;                    // radix class of the kernels to launch (fft_core.cuh)
  bool cube = false;                 // N == 4096: the 16x16x16 single-CTA kernels (kernels_cube.cuh)
  std::vector<float2> cube_tw0, cube_tw1;
  HostSubPlan s1, s2;                // s1: length N1 over stride-N2 columns; s2: length N2 over rows
  std::vector<float2> tw1, tw2;      // exp(-2 pi i k / F)
  std::vector<float2> twm;           // twm[p1*N2 + n2] = exp(-2 pi i k1(p1) n2 / N)       (forward: columns are natural n2)
  std::vector<float2> twm_inv;       // twm_inv[p1*N2 + p2] = exp(-2 pi i k1(p1) n2(p2) / N) (inverse: columns are positions);
                                     // empty when s2 is not a prime-factor transform (then equal to twm)
  // Coprime split (gcd(N1, N2) == 1): the four-step itself runs in Good-Thomas form. Time sample
  // n sits at (n1, n2) with n = (N2*n1 + N1*n2) mod N, frequency k at (k mod N1, k mod N2), and
  // W_N^(nk) = W_N1^(n1 k1) * W_N2^(n2 k2): a plain two-dimensional DFT, so the twiddle tables
  // above are not built and no kernel multiplies by them. Memory stays row-major in
  // (m, b) = (n div N2, n mod N2); the index maps below place a sample in its tile.
  bool gt = false;
  std::vector<int> fpos1;            // [N1] a = n mod N1 -> tile position of the length-N1 transform that takes the sample
  std::vector<int> fpos2;            // [N2] b = n mod N2 -> tile position of the length-N2 transform
  std::vector<int> col_lag;          // [N2] lag contribution of inverse column position q2: N1 * n2(q2) (non-gt: n2(q2))
};

inline int gcd_int(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

// Preferred coprime split of the lengths measured on B200 (N1 carries the radix-31 factor: the
// columns kernel fuses that butterfly with the magnitude / peak epilogue).
inline int measured_gt_n1(int N) {
  switch (N) {
    case 163680: return 341;         // 341 = 31*11, 480 = 15*32
    case 61380: return 279;          // 279 = 31*9, 220 = 11*20
    case 30690: return 341;          // 341 = 31*11, 90 = 9*10
    default: return 0;
  }
}

inline std::vector<int> prime_factors(int n) {
  std::vector<int> f;
  for (int p = 2; (long long)p * p <= n; ++p)
    while (n % p == 0) { f.push_back(p); n /= p; }
  if (n > 1) f.push_back(n);
  return f;
}

// Radix schedule for one tile transform: powers of two grouped into 16/8/4/2, odd primes
// as themselves; large odd radices first so the cheap power-of-two stages run at unit stride.
inline bool make_subplan(int F, HostSubPlan& sp, unsigned long long disabled = 0, const std::vector<int>* forced = nullptr) {
  sp = HostSubPlan();
  sp.F = F;
  if (F < 1) return false;
  for (int p : prime_factors(F))
    if (!radix_supported(p)) return false;
  // Fewest stages wins (each stage is one shared-memory pass and one barrier); among equals the
  // cheapest butterflies (smallest radix sum). Composite radices run in registers (fft_core.cuh).
  // (12, 20 and 22 also exist as in-register butterflies but measured slower on B200: register
  // pressure costs a resident CTA; they stay available through gnssacq_set_schedule.)
  static const int kStageRadices[] = {31, 25, 16, 15, 13, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2};
  std::vector<int> best, cur;
  std::function<void(int, size_t)> search = [&](int rem, size_t from) {
    if (rem == 1) {
      auto cost = [](const std::vector<int>& v) { int c = 0; for (int r : v) c += r; return c; };
      if (best.empty() || cur.size() < best.size() || (cur.size() == best.size() && cost(cur) < cost(best))) best = cur;
      return;
    }
    if (!best.empty() && cur.size() + 1 > best.size()) return;
    for (size_t a = from; a < sizeof(kStageRadices) / sizeof(int); ++a) {
      const int r = kStageRadices[a];
      if (rem % r) continue;
      if (r < 64 && ((disabled >> r) & 1ull)) continue;      // tuning: radix switched off
      cur.push_back(r);
      search(rem / r, a);              // non-increasing radices: largest butterflies first
      cur.pop_back();
    }
  };
  if (F > 1) search(F, 0);
  bool use_forced = false;
  // schedules measured faster than the rule above (tools/ab_sched.py, bench_configs.py)
  // 372 = 31 * 12: with the twiddle-free prime-factor form the in-register radix 12 = 4 x 3 saves a
  // shared-memory pass over 31 * 3 * 4 (columns kernel 24.7 -> 22.6 us per 32 units of 163680)
  // 480 = 15 * 32 and 352 = 11 * 32: two-stage prime-factor schedules with the in-register radix 32
  // (radix 32 is not part of the general search: it would re-plan the tuned power-of-two lengths)
  static const std::vector<std::vector<int>> kMeasured = {{10, 20}, {10, 25}, {11, 20}, {31, 12}, {15, 32}, {11, 32}};
  if (!disabled)
    for (const auto& m : kMeasured) {
      long long prod = 1;
      for (int r : m) prod *= r;
      if (prod == F) { best = m; use_forced = true; }
    }
  if (forced && !forced->empty()) {
    long long prod = 1;
    for (int r : *forced) prod *= r;
    if (prod == F) { best = *forced; use_forced = true; }
  }
  // order (measured, tools/ab_sched.py): warp-pair radices (31) first — the columns kernel fuses
  // stage 0 with its epilogue — then odd radices descending, then even ones descending
  // (440 = 11*5*8, 372 = 31*3*4, 320 = 5*8*8); pure powers of two ascending (128 = 8*16), which
  // gives the rows kernel's fused last stage the longer contiguous runs.
  if (!use_forced) {
    bool pow2_only = true;
    for (int r : best) pow2_only = pow2_only && (r & (r - 1)) == 0;
    std::sort(best.begin(), best.end(), [pow2_only](int a, int b) {
      if (pow2_only) return a < b;
      const int ca = a == 31 ? 0 : ((a & 1) ? 1 : 2), cb = b == 31 ? 0 : ((b & 1) ? 1 : 2);
      if (ca != cb) return ca < cb;
      return a > b;
    });
  }
  sp.radix = best;
  if (sp.radix.empty()) sp.radix.push_back(1);   // F == 1: no stage
  if (F == 1) sp.radix.clear();
  if ((int)sp.radix.size() > kMaxStages) return false;
  sp.m.resize(sp.radix.size());
  int m = F;
  for (size_t j = 0; j < sp.radix.size(); ++j) { m /= sp.radix[j]; sp.m[j] = m; }
  // frequency k = q0 + r0*(q1 + r1*(q2 + ...)) lands at position sum_j q_j * m_j
  sp.pos_of_freq.assign(F, 0);
  sp.freq_of_pos.assign(F, 0);
  sp.n_of_pos.resize(F);
  sp.pos_of_n.resize(F);
  for (int k = 0; k < F; ++k) { sp.n_of_pos[k] = k; sp.pos_of_n[k] = k; }
  for (int k = 0; k < F; ++k) {
    int rem = k, pos = 0;
    for (size_t j = 0; j < sp.radix.size(); ++j) { pos += (rem % sp.radix[j]) * sp.m[j]; rem /= sp.radix[j]; }
    sp.pos_of_freq[k] = pos;
    sp.freq_of_pos[pos] = k;
  }
  return true;
}

// W_F^k for k < F, then one butterfly-major table per stage (see SubPlan::tws_off).
inline std::vector<float2> unit_roots(HostSubPlan& sp) {
  const int F = sp.F;
  std::vector<float2> t(F);
  for (int k = 0; k < F; ++k) {
    double a = -2.0 * M_PI * (double)k / (double)F;
    t[k] = make_float2((float)cos(a), (float)sin(a));
  }
  sp.tws_off.assign(sp.radix.size(), 0);
  for (size_t j = 0; j < sp.radix.size(); ++j) {
    const int R = sp.radix[j], m = sp.m[j];
    if (m == 1) continue;
    sp.tws_off[j] = (int)t.size();
    for (int i = 0; i < m; ++i)
      for (int q = 1; q < R; ++q) {
        double a = -2.0 * M_PI * (double)((long long)q * i) / (double)(R * m);
        t.push_back(make_float2((float)cos(a), (float)sin(a)));
      }
  }
  if (!sp.radix.empty() && sp.m[0] > 1) {
    const int R = sp.radix[0], m = sp.m[0];
    sp.tws0_t_off = (int)t.size();
    for (int q = 1; q < R; ++q)
      for (int i = 0; i < m; ++i) {
        double a = -2.0 * M_PI * (double)((long long)q * i) / (double)(R * m);
        t.push_back(make_float2((float)cos(a), (float)sin(a)));
      }
  }
  return t;
}

// Whether make_plan accepts N (same conditions, no tables built).
inline bool plannable(int N) {
  if (N < 4) return false;
  for (int p : prime_factors(N))
    if (!radix_supported(p)) return false;
  int best = 0;
  for (int a = 1; (long long)a * a <= N; ++a)
    if (N % a == 0 && N / a <= kMaxSub) best = a;
  return best >= 2;
}

// Choose N = N1*N2 with both factors tile-sized and as square as possible.
// use_pfa(plan, which): whether the kernels that will run sub-transform `which` (1: length N1,
// 2: length N2) are the twiddle-free prime-factor ones (asked only for coprime schedules; the plan
// carries both radix schedules at that point, no tables yet).
inline bool make_plan(int N, HostPlan& pl, std::string& err, int force_n1 = 0, unsigned long long disabled = 0,
                      const std::vector<int>* sched1 = nullptr, const std::vector<int>* sched2 = nullptr,
                      const std::function<bool(const HostPlan&, int)>* use_pfa = nullptr, bool allow_gt = false) {
  pl = HostPlan();
  pl.N = N;
  if (N < 4) { err = "FFT length must be >= 4"; return false; }
  for (int p : prime_factors(N))
    if (!radix_supported(p)) {
      err = "FFT length " + std::to_string(N) + " has the unsupported prime factor " + std::to_string(p) +
            " (supported: 2, 3, 5, 7, 11, 13, 31)";
      return false;
    }
  int best = 0;
  for (int a = 1; (long long)a * a <= N; ++a)
    if (N % a == 0 && N / a <= kMaxSub) { best = a; }
  if (best == 0) { err = "FFT length " + std::to_string(N) + " cannot be split into two factors <= 1024"; return false; }
  // rows (contiguous, length N2) get the larger factor: longer coalesced runs — unless that
  // factor holds a warp-pair radix (31): the columns kernel fuses that butterfly with the
  // magnitude/peak epilogue, measured 5-9 % faster for 61380 = 279 x 220 and 30690 = 186 x 165.
  pl.N1 = best; pl.N2 = N / best;
  if (pl.N1 > 1 && pl.N2 % 31 == 0 && pl.N1 % 31 != 0) std::swap(pl.N1, pl.N2);
  if (allow_gt && force_n1 == 0 && N > kMidMax) {
    const int g = measured_gt_n1(N);
    if (g > 1 && N % g == 0 && gcd_int(g, N / g) == 1) { pl.N1 = g; pl.N2 = N / g; }
  }
  if (force_n1 > 1 && N % force_n1 == 0 && force_n1 <= kMaxSub && N / force_n1 <= kMaxSub && N / force_n1 > 1) {
    pl.N1 = force_n1; pl.N2 = N / force_n1;
  }
  if (pl.N1 < 2) { err = "FFT length " + std::to_string(N) + " is prime; unsupported"; return false; }
  pl.large = N > kMidMax;
  pl.cube = (N == 4096) && force_n1 == 0 && disabled == 0;
  if (pl.cube) {
    pl.cube_tw0.resize(15 * 256);
    pl.cube_tw1.resize(15 * 16);
    for (int q = 1; q < 16; ++q) {
      for (int t = 0; t < 256; ++t) {
        double a = -2.0 * M_PI * (double)(q * t) / 4096.0;
        pl.cube_tw0[(q - 1) * 256 + t] = make_float2((float)cos(a), (float)sin(a));
      }
      for (int c = 0; c < 16; ++c) {
        double a = -2.0 * M_PI * (double)(q * c) / 256.0;
        pl.cube_tw1[(q - 1) * 16 + c] = make_float2((float)cos(a), (float)sin(a));
      }
    }
  }
  if (!make_subplan(pl.N1, pl.s1, disabled, sched1) || !make_subplan(pl.N2, pl.s2, disabled, sched2)) { err = "unsupported factorisation"; return false; }
  pl.gt = allow_gt && pl.large && gcd_int(pl.N1, pl.N2) == 1;
  if (use_pfa && pl.large) {
    if (coprime_schedule(pl.s1.radix) && (*use_pfa)(pl, 1)) set_pfa(pl.s1);
    if (coprime_schedule(pl.s2.radix) && (*use_pfa)(pl, 2)) set_pfa(pl.s2);
  }
  for (int r : pl.s1.radix) pl.rclass = std::max(pl.rclass, radix_class_of(r));
  for (int r : pl.s2.radix) pl.rclass = std::max(pl.rclass, radix_class_of(r));
  pl.tw1 = unit_roots(pl.s1);
  pl.tw2 = unit_roots(pl.s2);
  pl.col_lag.resize(pl.N2);
  for (int q2 = 0; q2 < pl.N2; ++q2) pl.col_lag[q2] = (pl.gt ? pl.N1 : 1) * pl.s2.n_of_pos[q2];
  if (pl.gt) {
    const int inv2 = modinv(pl.N2 % pl.N1, pl.N1), inv1 = modinv(pl.N1 % pl.N2, pl.N2);
    pl.fpos1.resize(pl.N1);
    pl.fpos2.resize(pl.N2);
    for (int a = 0; a < pl.N1; ++a) pl.fpos1[a] = pl.s1.pos_of_n[(int)(((long long)a * inv2) % pl.N1)];
    for (int b = 0; b < pl.N2; ++b) pl.fpos2[b] = pl.s2.pos_of_n[(int)(((long long)b * inv1) % pl.N2)];
    pl.twm.assign(1, make_float2(1.f, 0.f));       // never read
    return true;
  }
  pl.twm.resize((size_t)N);
  for (int p1 = 0; p1 < pl.N1; ++p1) {
    const long long k1 = pl.s1.freq_of_pos[p1];
    for (int n2 = 0; n2 < pl.N2; ++n2) {
      long long e = (k1 * n2) % N;
      double a = -2.0 * M_PI * (double)e / (double)N;
      pl.twm[(size_t)p1 * pl.N2 + n2] = make_float2((float)cos(a), (float)sin(a));
    }
  }
  if (pl.s2.pfa) {
    pl.twm_inv.resize((size_t)N);
    for (int p1 = 0; p1 < pl.N1; ++p1)
      for (int p2 = 0; p2 < pl.N2; ++p2)
        pl.twm_inv[(size_t)p1 * pl.N2 + p2] = pl.twm[(size_t)p1 * pl.N2 + pl.s2.n_of_pos[p2]];
  }
  return true;
}

}  // namespace acq
