// libgnssacq.so — C ABI (include/gnssacq.h) over the sm_100a acquisition kernels.
#include "../../include/gnssacq.h"
#include "fft_plan.h"
#include "kernels.cuh"
#include "registry.h"
#include "preprocess.cuh"
#include "kernels_cube.cuh"
#include "bank.cuh"
#include "async_copy.cuh"
#ifndef GNSSACQ_EMU_BUILD
#include <cuda.h>
#endif

#include <dlfcn.h>

#include <algorithm>
#include <functional>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

using namespace acq;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(GNSSACQ_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
  } while (0)

// Working-set budgets. Units are ordered replica-fastest, so a Doppler bin's spectra X[d] are
// re-read by R consecutive units and then never again: the capture-spectra chunk may be large
// (memory-bound only). What must stay L2-resident (B200 L2 ~126 MB) is the replica spectra C
// plus the inverse-FFT scratch between the rows and columns kernels.
constexpr size_t kXChunkBytes = 512u << 20;
constexpr size_t kScratchBytes = 40u << 20;     // split over the two lanes when chunks overlap

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { p = nullptr; return fail(GNSSACQ_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
    cap = bytes;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

}  // namespace

struct gnssacq {
  static constexpr int kMaxLanes = 4;
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  int64_t launches = 0;
  size_t smem_optin = 0;
  int num_sms = 148;

  std::vector<double> nco_c128;       // 1024 x (re, im)
  DevBuf d_nco_f32, d_nco_f64;

  DevBuf d_x_own;                     // capture (owned copy)
  const float2* d_x = nullptr;
  int64_t n_x = 0;

  HostPlan hp;
  DevPlan dp{};
  DevBuf d_tw1, d_tw2, d_twm, d_twm_inv, d_maps, d_cube0, d_cube1;
  CubeTw cube_tw{nullptr, nullptr};
  DevBuf d_C;                         // replica spectra [R][N]
  int R = 0, N = 0;
  void* nccl_comm = nullptr;          // ncclComm_t of gnssacq_nccl_init (NCCL is bound at run time, see NcclApi)
  int nccl_rank = 0, nccl_world = 1;
  DevBuf d_allrec, d_merged;
  int embed_n = 0;                    // != 0: the caller's transform length, embedded in the planned length N (>= 2*embed_n - 1)

  DevBuf d_X, d_scratch, d_parts, d_freq, d_rec, d_q, d_tmp;
  DevBuf d_raw, d_ext, d_y1, d_z, d_fir, d_pre128;        // capture front end
  DevBuf d_chips, d_base, d_bank;                         // replica builder / correlator bank
  DevBuf d_repext;                                        // periodically extended replicas of an embedded length

  // optional per-stage timing (gnssacq_set_profiling): event pairs recorded around the
  // launches of each stage, folded into prof_ms at gnssacq_get_stage_times().
  bool use_spec = true;               // plan-specialised correlate kernels when one matches
  bool use_gt = true;                 // coprime four-step splits (no twiddle pass) for the lengths that have one
  // copy-engine-fed correlate pair (kernels_v3.cuh) for coprime plans with two-stage schedules
  bool use_v3 = true;
  int v3_rows_variant = 0, v3_cols_variant = 0;      // tile shapes (registry.cu), A/B
  int v3_rc = 0, v3_g = 0;                           // replicas x Doppler bins per launch (0 = automatic)
  bool fwd_v6 = true;                                // forward rows pass through k_fwd_rows_v6 where instantiated
  DevBuf d_v3tab;                                    // padded column table + tile origins
  // one persistent kernel per Doppler chunk (kernels_fused.cuh) instead of the pair: correct, measured 14 %
  // slower than the pair on config 2 (profiles/README.md r03f), hence off unless asked for
  bool use_fused = false;
  int fused_rc = 0, fused_g = 0, fused_sets = 3, fused_ctas = 0, fused_tpt = 0;   // group shape, scratch ring depth, CTAs per SM (0 = automatic)
  DevBuf d_fsync;                                    // ticket, error flag, per-group counters
  DevBuf d_hint;                                     // per (replica, Doppler) unit: best value reported so far (peak-search floor)
  int v3tab_key[4] = {0, 0, 0, 0};                   // (N, RB, PB, CW) the table was built for
  int v3_ntiles = 0;
  struct LaneMap { TensorMap map; const void* base = nullptr; long long slots = 0; int NP = 0, CW = 0, F1 = 0, F2 = 0; } v3_map[kMaxLanes + 1];
  int small_ctas = 3;                 // bit 0: 8-row / 128-160-thread rows kernel, bit 1: 128-thread columns kernel
  int force_n1 = 0;                   // tuning: force the four-step split N = n1 * (N/n1)
  unsigned long long disabled_radices = 0;   // tuning: stage radices the planner may not use
  bool plan_dirty = false;
  std::vector<int> forced_sched[2];          // tuning: explicit stage lists for the N1 / N2 transforms
  int force_uc = 0;                   // tuning: force the number of units per correlate launch
  bool overlap = true;                // large plans: alternate unit chunks over two streams so the
                                      // rows kernel of one chunk overlaps the columns kernel of the other
  int nlanes = 2;
  size_t scratch_bytes = kScratchBytes;   // total over all lanes
  size_t xchunk_bytes = kXChunkBytes;
  cudaStream_t lane[kMaxLanes] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxLanes] = {};
  // gnssacq_set_signal copies on its own stream, so that the capture's host-to-device transfer overlaps the replica set-up
  // that usually follows it; whoever reads or writes the capture buffer (or synchronises for the host) joins it first
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_x = nullptr, ev_before_x = nullptr;
  bool x_pending = false;
  DevBuf d_scratch_lane[kMaxLanes];
  bool profiling = false;
  struct Span { int stage; cudaEvent_t a, b; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> event_pool;
  double prof_ms[4] = {0, 0, 0, 0};
  int64_t prof_launches[4] = {0, 0, 0, 0};
};

namespace {

// plan-specialised kernels: off for embedded lengths, whose zero-padded blocks and partial sums only
// the generic kernels implement
bool spec_on(const gnssacq* h) { return h->use_spec && h->embed_n == 0; }

// The handle's stream waits for a capture copy that is still on the copy stream (gnssacq_set_signal).
int join_capture_copy(gnssacq* h) {
  if (h->x_pending) {
    CU(cudaStreamWaitEvent(h->stream, h->ev_x, 0));
    h->x_pending = false;
  }
  return 0;
}
// Every synchronisation for the host covers that copy too (include/gnssacq.h: a pinned capture buffer is the caller's
// again after any call that synchronises).
#define SYNC_MAIN(h)                                             \
  do {                                                           \
    if (int rc_ = join_capture_copy(h)) return rc_;              \
    CU(cudaStreamSynchronize((h)->stream));                      \
  } while (0)

int upload_nco(gnssacq* h) {
  std::vector<float2> f32(kNcoSize);
  for (int k = 0; k < kNcoSize; ++k)
    f32[k] = make_float2((float)h->nco_c128[2 * k], (float)h->nco_c128[2 * k + 1]);
  if (int rc = h->d_nco_f32.ensure(kNcoSize * sizeof(float2))) return rc;
  if (int rc = h->d_nco_f64.ensure(kNcoSize * sizeof(double2))) return rc;
  CU(cudaMemcpyAsync(h->d_nco_f32.p, f32.data(), kNcoSize * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->d_nco_f64.p, h->nco_c128.data(), kNcoSize * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  SYNC_MAIN(h);
  return 0;
}

void fill_subplan(const HostSubPlan& hs, const float2* tw, SubPlan& sp) {
  sp.F = hs.F;
  sp.ns = (int)hs.radix.size();
  for (int j = 0; j < kMaxStages; ++j) {
    sp.radix[j] = j < sp.ns ? hs.radix[j] : 1;
    sp.m[j] = j < sp.ns ? hs.m[j] : 1;
    sp.tws_off[j] = j < sp.ns ? hs.tws_off[j] : 0;
  }
  sp.tws0_t_off = hs.tws0_t_off;
  sp.tw = tw;
  sp.pfa = hs.pfa ? 1 : 0;
}

int upload_plan(gnssacq* h, int N) {
  if (h->hp.N == N && !h->plan_dirty) return 0;
  HostPlan hp;
  std::string err;
  // A coprime radix schedule runs as a twiddle-free prime-factor transform when every kernel that
  // touches that sub-transform is a plan-specialised one (they carry the index maps; the generic
  // runtime-planned kernels are Cooley-Tukey only).
  const bool spec = spec_on(h);
  const std::function<bool(const HostPlan&, int)> use_pfa = [spec](const HostPlan& hpl, int which) {
    if (!spec) return false;
    auto schedule_of = [](const HostSubPlan& hs, bool pfa) {      // schedule only: the twiddle tables do not exist yet
      SubPlan sp{};
      sp.F = hs.F;
      sp.ns = (int)hs.radix.size();
      for (int j = 0; j < kMaxStages; ++j) {
        sp.radix[j] = j < sp.ns ? hs.radix[j] : 1;
        sp.m[j] = j < sp.ns ? hs.m[j] : 1;
      }
      sp.pfa = pfa ? 1 : 0;
      return sp;
    };
    if (which == 1) {
      const SubPlan s1 = schedule_of(hpl.s1, true);
      return find_cols_kernel(s1, false) && find_cols_kernel(s1, true) && find_fwd_cols_kernel(s1, 0) && find_fwd_cols_kernel(s1, 1);
    }
    // The columns of the inverse are tile positions of the length-N2 transform: only the specialised
    // columns kernels translate them back to lags (n2_of_pos), so they must exist for s1 as it will
    // run (prime-factor or not — s1 has been decided by now).
    const SubPlan s1 = schedule_of(hpl.s1, hpl.s1.pfa);
    const SubPlan s2 = schedule_of(hpl.s2, true);
    return find_rows_kernel(s2, hpl.gt) && find_fwd_rows_kernel(s2) && find_cols_kernel(s1, false) && find_cols_kernel(s1, true);
  };
  if (!make_plan(N, hp, err, h->force_n1, h->disabled_radices, &h->forced_sched[0], &h->forced_sched[1], &use_pfa, h->use_gt)) return fail(GNSSACQ_EINVAL, err);
  h->plan_dirty = false;
  if (int rc = h->d_tw1.ensure(hp.tw1.size() * sizeof(float2))) return rc;
  if (int rc = h->d_tw2.ensure(hp.tw2.size() * sizeof(float2))) return rc;
  if (int rc = h->d_twm.ensure(hp.twm.size() * sizeof(float2))) return rc;
  CU(cudaMemcpyAsync(h->d_tw1.p, hp.tw1.data(), hp.tw1.size() * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->d_tw2.p, hp.tw2.data(), hp.tw2.size() * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->d_twm.p, hp.twm.data(), hp.twm.size() * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
  if (!hp.twm_inv.empty()) {
    if (int rc = h->d_twm_inv.ensure(hp.twm_inv.size() * sizeof(float2))) return rc;
    CU(cudaMemcpyAsync(h->d_twm_inv.p, hp.twm_inv.data(), hp.twm_inv.size() * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
  }
  // index maps: n1_of_pos [N1], n2_of_pos [N2], pos2_of_n [N2]
  std::vector<int> maps;
  maps.insert(maps.end(), hp.s1.n_of_pos.begin(), hp.s1.n_of_pos.end());
  maps.insert(maps.end(), hp.s2.n_of_pos.begin(), hp.s2.n_of_pos.end());
  maps.insert(maps.end(), hp.s2.pos_of_n.begin(), hp.s2.pos_of_n.end());
  maps.insert(maps.end(), hp.col_lag.begin(), hp.col_lag.end());
  maps.insert(maps.end(), hp.fpos1.begin(), hp.fpos1.end());          // empty unless hp.gt
  maps.insert(maps.end(), hp.fpos2.begin(), hp.fpos2.end());
  if (int rc = h->d_maps.ensure(maps.size() * sizeof(int))) return rc;
  CU(cudaMemcpyAsync(h->d_maps.p, maps.data(), maps.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  if (hp.cube) {
    if (int rc = h->d_cube0.ensure(hp.cube_tw0.size() * sizeof(float2))) return rc;
    if (int rc = h->d_cube1.ensure(hp.cube_tw1.size() * sizeof(float2))) return rc;
    CU(cudaMemcpyAsync(h->d_cube0.p, hp.cube_tw0.data(), hp.cube_tw0.size() * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_cube1.p, hp.cube_tw1.data(), hp.cube_tw1.size() * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
    h->cube_tw.tw0 = h->d_cube0.as<float2>();
    h->cube_tw.tw1 = h->d_cube1.as<float2>();
  }
  SYNC_MAIN(h);
  h->dp.N = hp.N; h->dp.N1 = hp.N1; h->dp.N2 = hp.N2;
  fill_subplan(hp.s1, h->d_tw1.as<float2>(), h->dp.s1);
  fill_subplan(hp.s2, h->d_tw2.as<float2>(), h->dp.s2);
  h->dp.twm = h->d_twm.as<float2>();
  h->dp.twm_inv = hp.twm_inv.empty() ? h->d_twm.as<float2>() : h->d_twm_inv.as<float2>();
  h->dp.n1_of_pos = h->d_maps.as<int>();
  h->dp.n2_of_pos = h->d_maps.as<int>() + hp.N1;
  h->dp.pos2_of_n = h->d_maps.as<int>() + hp.N1 + hp.N2;
  h->dp.col_lag = h->d_maps.as<int>() + hp.N1 + 2 * hp.N2;
  h->dp.gt = hp.gt ? 1 : 0;
  h->dp.xlen = h->embed_n ? h->embed_n : hp.N;
  h->dp.sum_lags = h->dp.xlen;
  h->dp.fpos1 = h->d_maps.as<int>() + hp.N1 + 3 * hp.N2;
  h->dp.fpos2 = h->dp.fpos1 + hp.N1;
  h->hp = std::move(hp);
  h->v3tab_key[0] = 0;                               // tables derived from the previous plan are stale
  return 0;
}

size_t mid_smem(const DevPlan& p, bool with_q) {
  return (size_t)(p.N1 * mid_sa(p.N2) + p.N2 * mid_sb(p.N1)) * sizeof(float2) + (with_q ? (size_t)p.N * sizeof(float) : 0);
}
size_t cols_smem(const DevPlan& p, bool with_q) {
  return (size_t)p.N1 * kTileW * (sizeof(float2) + (with_q ? sizeof(float) : 0));
}
size_t rows_smem(const DevPlan& p) { return (size_t)p.N2 * kRowPitch * sizeof(float2); }

template <class K> int allow_smem(gnssacq* h, K kern, size_t bytes) {
  if (bytes > h->smem_optin) return fail(GNSSACQ_EINVAL, "transform does not fit in shared memory");
  // static shared memory (reduction scratch) counts against the 48 KB default too
  if (bytes > 40 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

enum { kStageFwd = 0, kStageCorrRows = 1, kStageCorrCols = 2, kStageFinalize = 3 };

cudaEvent_t take_event(gnssacq* h) {
  if (!h->event_pool.empty()) { cudaEvent_t e = h->event_pool.back(); h->event_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
struct StageTimer {
  gnssacq* h; int stage; int nl; cudaEvent_t a = nullptr;
  StageTimer(gnssacq* h_, int stage_, int nl_) : h(h_), stage(stage_), nl(nl_) {
    if (h->profiling) { a = take_event(h); cudaEventRecord(a, h->stream); }
  }
  ~StageTimer() {
    if (a) {
      cudaEvent_t b = take_event(h);
      cudaEventRecord(b, h->stream);
      h->spans.push_back({stage, a, b});
      h->prof_launches[stage] += nl;
    }
  }
};

// Run f(integral_constant<RC>) for the plan's radix class.
template <class F> int with_radix_class(int rc, F&& f) {
  switch (rc) {
    case 0: return f(std::integral_constant<int, 0>{});
    case 1: return f(std::integral_constant<int, 1>{});
    default: return f(std::integral_constant<int, 2>{});
  }
}

// Forward transforms of `nt` inputs into X (position order). SRC 0: capture blocks with
// wipe-off (nt = Dc*B, transform d*B+b); SRC 1: real replicas.
template <int SRC>
int forward(gnssacq* h, const float* rep, const double* d_freq, int stride, int B, int nt, float2* X) {
  if (SRC == 0)
    if (int rc = join_capture_copy(h)) return rc;              // the capture may still be on its way (gnssacq_set_signal)
  return with_radix_class(h->hp.rclass, [&](auto rc) -> int {
    constexpr int RC = decltype(rc)::value;
    const DevPlan& p = h->dp;
    const float2* tab = h->d_nco_f32.as<float2>();
    StageTimer timer(h, kStageFwd, h->hp.large ? 2 : 1);
    if (h->hp.cube && spec_on(h)) {
      auto kern = k_fwd_cube<SRC>;
      GNSSACQ_LAUNCH(kern, dim3(nt), dim3(256), (size_t)kCubeSmem, h->stream, h->cube_tw, h->d_x, rep, d_freq, tab, stride, B, X);
      h->launches += 1;
    } else if (!h->hp.large) {
      const size_t sm = mid_smem(p, false);
      auto kern = k_fwd_mid<RC, SRC>;
      if (int rc2 = allow_smem(h, kern, sm)) return rc2;
      GNSSACQ_LAUNCH(kern, dim3(nt), dim3(kThreads), sm, h->stream, p, h->d_x, rep, d_freq, tab, stride, B, X);
      h->launches += 1;
    } else {
      const size_t smc = cols_smem(p, false), smr = rows_smem(p);
      fwd_cols_fn kc = spec_on(h) ? find_fwd_cols_kernel(p.s1, SRC) : nullptr;
      fwd_rows_fn kr = spec_on(h) ? find_fwd_rows_kernel(p.s2) : nullptr;
      if (!kc) kc = k_fwd_cols<RC, SRC>;
      if (!kr) kr = k_fwd_rows<RC>;
      if (int rc2 = allow_smem(h, kc, smc)) return rc2;
      if (int rc2 = allow_smem(h, kr, smr)) return rc2;
      GNSSACQ_LAUNCH(kc, dim3((p.N2 + kTileW - 1) / kTileW, nt), dim3(kThreads), smc, h->stream,
                     p, h->d_x, rep, d_freq, tab, stride, B, X);
      // rows pass: the copy-engine-fed two-role kernel for coprime plans that have one ("fwd_v6", default on)
      const FwdRowsV6 f6 = (spec_on(h) && h->use_v3 && h->fwd_v6 && p.gt) ? find_fwd_rows_v6(p.s2) : FwdRowsV6{nullptr, 0, 0, 0, 0};
      if (f6.fn && f6.smem <= h->smem_optin) {
        if (int rc2 = allow_smem(h, f6.fn, f6.smem)) return rc2;
        const int nrt = (p.N1 + f6.T - 1) / f6.T;
        // one wave of CTAs (equal work each); a CTA walks its share of the transforms with the rows of its tile
        const int split = std::max(1, std::min(nt, h->num_sms * f6.ctas_per_sm / nrt));
        GNSSACQ_LAUNCH(f6.fn, dim3(nrt, split), dim3(f6.threads), f6.smem, h->stream, p, X, nt);
      } else {
        GNSSACQ_LAUNCH(kr, dim3((p.N1 + kTileW - 1) / kTileW, nt), dim3(kThreads), smr, h->stream, p, X);
      }
      h->launches += 2;
    }
    CU(cudaGetLastError());
    return 0;
  });
}

// ---------------------------------------------------------------- copy-engine-fed pair (kernels_v3.cuh)
struct V3Setup {
  bool on = false;
  RowsV3 r{};
  ColsV3 c{};
  FusedKernel f{};                  // f.fn != nullptr: the fused persistent kernel runs instead of the pair
  int ntiles = 0, NP = 0, F1 = 0, F2 = 0;
};

// The pair runs when the plan is a coprime split and both tile transforms have an instantiation.
V3Setup v3_setup(const gnssacq* h, bool multi, bool dump) {
  V3Setup v;
  if (!h->use_v3 || !spec_on(h) || !h->hp.large || !h->hp.gt || !h->hp.s1.pfa || !h->hp.s2.pfa) return v;
  v.r = find_rows_v3(h->dp.s2, h->v3_rows_variant);
  v.c = find_cols_v3(h->dp.s1, multi, dump, h->v3_cols_variant);
  if (!v.r.fn || !v.c.fn) return v;
  if (v.r.smem > h->smem_optin || v.c.smem > h->smem_optin) return v;
  if (h->use_fused) {
    v.f = find_fused(h->dp.s1, h->dp.s2, multi, dump);
    if (v.f.fn && v.f.smem <= h->smem_optin) {          // the fused kernel fixes the tile shapes
      v.r.T = v.f.T; v.r.RA = v.f.RA; v.r.RB = v.f.RB; v.r.PB = v.f.PB;
      v.c.CW = v.f.CW;
    } else v.f = FusedKernel{};
  }
  v.NP = v.r.RA * v.r.PB;
  // ragged layouts walk the flat padded row; its last PB - RB columns are pads, a tile of nothing but those is dropped
  v.ntiles = (v.r.RB % v.c.CW == 0) ? v.r.RA * (v.r.RB / v.c.CW) : (v.NP - (v.r.PB - v.r.RB) + v.c.CW - 1) / v.c.CW;
  // tensor-map boxes are limited to 256 per dimension: N1 = F1 * F2 with the largest F1 <= 256
  for (int f = 1; f <= 256 && f <= h->hp.N1; ++f)
    if (h->hp.N1 % f == 0) v.F1 = f;
  v.F2 = h->hp.N1 / v.F1;
  if (v.F2 > 256) return v;
  v.on = true;
  return v;
}

// Padded column table (lag term of scratch column a'*PB + b', -1 for pad columns, -1 slack behind
// the row for partly out-of-range tiles) followed by the tile origins.
int v3_upload_tables(gnssacq* h, const V3Setup& v) {
  const int key[4] = {h->hp.N, v.r.RB, v.r.PB, v.c.CW};
  if (h->d_v3tab.p && std::equal(key, key + 4, h->v3tab_key) && h->v3_ntiles == v.ntiles) return 0;
  const int ncol = v.NP + 2 * v.c.CW;
  std::vector<int> tab((size_t)ncol + v.ntiles, -1);
  for (int a = 0; a < v.r.RA; ++a)
    for (int b = 0; b < v.r.RB; ++b) tab[(size_t)a * v.r.PB + b] = h->hp.col_lag[(size_t)a * v.r.RB + b];
  const bool exact = v.r.RB % v.c.CW == 0;
  const int per = exact ? v.r.RB / v.c.CW : 0;
  for (int t = 0; t < v.ntiles; ++t) tab[(size_t)ncol + t] = exact ? (t / per) * v.r.PB + (t % per) * v.c.CW : t * v.c.CW;
  if (int rc = h->d_v3tab.ensure(tab.size() * sizeof(int))) return rc;
  CU(cudaMemcpyAsync(h->d_v3tab.p, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  SYNC_MAIN(h);            // `tab` is a local
  std::copy(key, key + 4, h->v3tab_key);
  h->v3_ntiles = v.ntiles;
  return 0;
}

int v3_map_for(gnssacq* h, const V3Setup& v, int which, const void* base, long long slots) {
  gnssacq::LaneMap& m = h->v3_map[which];
  if (m.base == base && m.slots == slots && m.NP == v.NP && m.CW == v.c.CW && m.F1 == v.F1 && m.F2 == v.F2) return 0;
  const unsigned long long dims[3] = {(unsigned long long)v.NP, (unsigned long long)v.F1, (unsigned long long)v.F2 * (unsigned long long)slots};
  const unsigned long long strides[2] = {(unsigned long long)v.NP * sizeof(float2), (unsigned long long)v.NP * sizeof(float2) * v.F1};
  const unsigned box[3] = {(unsigned)v.c.CW, (unsigned)v.F1, (unsigned)v.F2};
  const int rc = encode_tensor_map_3d_u64(&m.map, base, dims, strides, box);
  if (rc != 0) return fail(GNSSACQ_ECUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string(rc));
  m.base = base; m.slots = slots; m.NP = v.NP; m.CW = v.c.CW; m.F1 = v.F1; m.F2 = v.F2;
  return 0;
}

// Chunk shape of the pair: replicas x Doppler bins per launch.
void v3_chunk_shape(const gnssacq* h, int B, int dc, int& Rc, int& G) {
  G = h->v3_g > 0 ? h->v3_g : std::max(1, 8 / B);    // measured on config 2 (tools/ab_v3.py): 16 x 8 units per launch
  G = std::max(1, std::min(G, dc));
  Rc = h->v3_rc > 0 ? h->v3_rc : 16;                 // measured on config 2 (tools/ab_v3.py): 16 x 4 units per launch
  Rc = std::max(1, std::min(Rc, h->R));
}

int correlate_chunk_v3(gnssacq* h, const V3Setup& v, int B, int D, int d0, int dc, int n_lags, float scale, float* d_qdump) {
  const DevPlan& p = h->dp;
  const int R = h->R;
  int Rc, G;
  v3_chunk_shape(h, B, dc, Rc, G);
  const long long slots = (long long)Rc * G * B;
  const size_t sbytes = (size_t)slots * p.N1 * v.NP * sizeof(float2);
  const int nchunks = ((dc + G - 1) / G) * ((R + Rc - 1) / Rc);
  const bool two_lanes = h->overlap && nchunks > 1;
  if (int rc = v3_upload_tables(h, v)) return rc;
  if (two_lanes) {
    for (int l = 0; l < h->nlanes; ++l) {
      if (int rc = h->d_scratch_lane[l].ensure(sbytes)) return rc;
      if (int rc = v3_map_for(h, v, l, h->d_scratch_lane[l].p, slots)) return rc;
    }
  } else {
    if (int rc = h->d_scratch.ensure(sbytes)) return rc;
    if (int rc = v3_map_for(h, v, gnssacq::kMaxLanes, h->d_scratch.p, slots)) return rc;
  }
  if (int rc = allow_smem(h, v.r.fn, v.r.smem)) return rc;
  if (int rc = allow_smem(h, v.c.fn, v.c.smem)) return rc;
  DevPlan pv = p;                                    // the columns kernel indexes the padded column table
  pv.col_lag = h->d_v3tab.as<int>();
  const int* tile_col0 = h->d_v3tab.as<int>() + v.NP + 2 * v.c.CW;
  const int nrt = (p.N1 + v.r.T - 1) / v.r.T;
  const int rows_slots = h->num_sms * v.r.ctas_per_sm, cols_slots = h->num_sms * v.c.ctas_per_sm;
  StageTimer timer(h, kStageCorrCols, 0);
  if (two_lanes) {
    CU(cudaEventRecord(h->ev_fork, h->stream));
    for (int l = 0; l < h->nlanes; ++l) CU(cudaStreamWaitEvent(h->lane[l], h->ev_fork, 0));
  }
  int k = 0, nl = 0;
  for (int dd0 = 0; dd0 < dc; dd0 += G)
    for (int r0 = 0; r0 < R; r0 += Rc, ++k) {
      ChunkV3 ck{r0, std::min(Rc, R - r0), dd0, std::min(G, dc - dd0)};
      const int l = two_lanes ? k % h->nlanes : gnssacq::kMaxLanes;
      cudaStream_t st = two_lanes ? h->lane[l] : h->stream;
      float2* scr = two_lanes ? h->d_scratch_lane[l].as<float2>() : h->d_scratch.as<float2>();
      const int pairs = ck.G * B;
      int split = 1;                                  // split the (Doppler, block) list when the grid would not fill the GPU twice
      if (nrt * ck.Rc < 2 * rows_slots) split = std::min(pairs, (2 * rows_slots + nrt * ck.Rc - 1) / (nrt * ck.Rc));
      GNSSACQ_LAUNCH(v.r.fn, dim3(nrt, ck.Rc, split), dim3(v.r.threads), v.r.smem, st, p, h->d_X.as<float2>(), h->d_C.as<float2>(), ck, B, scr);
      const int ntasks = ck.Rc * ck.G * v.ntiles;
      GNSSACQ_LAUNCH(v.c.fn, dim3(std::min(ntasks, cols_slots)), dim3(v.c.threads), v.c.smem, st, pv, h->v3_map[l].map, v.F2, tile_col0,
                     ck, B, D, d0, n_lags, scale, v.ntiles, h->d_parts.as<Part>(), d_qdump, h->d_hint.as<unsigned>());
      nl += 2;
    }
  h->launches += nl;
  h->prof_launches[kStageCorrCols] += h->profiling ? nl : 0;
  if (two_lanes)
    for (int l = 0; l < h->nlanes; ++l) {
      CU(cudaEventRecord(h->ev_join[l], h->lane[l]));
      CU(cudaStreamWaitEvent(h->stream, h->ev_join[l], 0));
    }
  CU(cudaGetLastError());
  return 0;
}

// The whole correlate stage of a Doppler chunk as one persistent launch (kernels_fused.cuh).
int correlate_fused(gnssacq* h, const V3Setup& v, int B, int D, int d0, int dc, int n_lags, float scale, float* d_qdump) {
  const DevPlan& p = h->dp;
  FusedJob job{};
  job.R = h->R; job.dc = dc; job.B = B;
  // group = Rc replicas x G bins x B blocks; about 16 unit-blocks per scratch set keeps three sets (and the
  // replica spectra) inside L2
  job.G = h->fused_g > 0 ? h->fused_g : std::max(1, 4 / B);
  job.G = std::max(1, std::min(job.G, dc));
  job.Rc = h->fused_rc > 0 ? h->fused_rc : std::max(4, 16 / (job.G * B));
  job.Rc = std::max(1, std::min(job.Rc, h->R));
  job.ngr = (h->R + job.Rc - 1) / job.Rc;
  job.ng = job.ngr * ((dc + job.G - 1) / job.G);
  job.nrt = (p.N1 + v.f.T - 1) / v.f.T;
  job.ntiles = v.ntiles;
  job.nsets = std::max(2, std::min(h->fused_sets, 8));
  job.slots_per_set = job.Rc * job.G * B;
  job.nR = job.nrt * job.Rc;
  job.tpt = std::max(1, std::min(h->fused_tpt > 0 ? h->fused_tpt : 5, job.ntiles));
  job.nC = ((job.ntiles + job.tpt - 1) / job.tpt) * job.Rc * job.G;
  job.D = D; job.d0 = d0; job.n_lags = n_lags; job.zmul = v.F2; job.scale = scale;
  if ((long long)job.ng * (job.nR + job.nC) > 0x7fffffffll) return fail(GNSSACQ_EINVAL, "too many tasks for one fused launch");
  const long long slots = (long long)job.nsets * job.slots_per_set;
  const size_t sbytes = (size_t)slots * p.N1 * v.NP * sizeof(float2);
  if (int rc = v3_upload_tables(h, v)) return rc;
  if (int rc = h->d_scratch.ensure(sbytes)) return rc;
  if (int rc = v3_map_for(h, v, gnssacq::kMaxLanes, h->d_scratch.p, slots)) return rc;
  const size_t nsync = 2 + 2 * (size_t)job.ng;
  if (int rc = h->d_fsync.ensure(nsync * sizeof(int))) return rc;
  CU(cudaMemsetAsync(h->d_fsync.p, 0, nsync * sizeof(int), h->stream));
  FusedSync sy;
  sy.ticket = h->d_fsync.as<int>();
  sy.error = sy.ticket + 1;
  sy.rows_done = sy.ticket + 2;
  sy.cols_done = sy.rows_done + job.ng;
  if (int rc = allow_smem(h, v.f.fn, v.f.smem)) return rc;
  DevPlan pv = p;
  pv.col_lag = h->d_v3tab.as<int>();
  const int* tile_col0 = h->d_v3tab.as<int>() + v.NP + 2 * v.c.CW;
  const int ctas = h->fused_ctas > 0 ? h->fused_ctas : v.f.ctas_per_sm;
  const long long total = (long long)job.ng * (job.nR + job.nC);
  const int grid = (int)std::min<long long>(total, (long long)h->num_sms * ctas);
  StageTimer timer(h, kStageCorrCols, 1);
  GNSSACQ_LAUNCH(v.f.fn, dim3(grid), dim3(v.f.threads), v.f.smem, h->stream, pv, h->v3_map[gnssacq::kMaxLanes].map, tile_col0, job, sy,
                 h->d_X.as<float2>(), h->d_C.as<float2>(), h->d_scratch.as<float2>(), h->d_parts.as<Part>(), d_qdump, h->d_hint.as<unsigned>());
  h->launches += 1;
  CU(cudaGetLastError());
  return 0;
}

// Correlate one doppler chunk [d0, d0+dc) against all replicas.
int correlate_chunk(gnssacq* h, int B, int D, int d0, int dc, int Uc, int n_lags, float scale, int ntiles, float* d_qdump) {
  return with_radix_class(h->hp.rclass, [&](auto rc) -> int {
    constexpr int RC = decltype(rc)::value;
    const DevPlan& p = h->dp;
    const int R = h->R;
    if (h->hp.cube && spec_on(h)) {
      StageTimer timer(h, kStageCorrCols, 1);
      if (B > 1)
        GNSSACQ_LAUNCH(k_corr_cube<true>, dim3(R * dc), dim3(256), (size_t)kCubeSmem, h->stream, h->cube_tw, h->d_X.as<float2>(),
                       h->d_C.as<float2>(), R, B, D, d0, n_lags, scale, h->d_parts.as<Part>(), d_qdump);
      else
        GNSSACQ_LAUNCH(k_corr_cube<false>, dim3(R * dc), dim3(256), (size_t)kCubeSmem, h->stream, h->cube_tw, h->d_X.as<float2>(),
                       h->d_C.as<float2>(), R, B, D, d0, n_lags, scale, h->d_parts.as<Part>(), d_qdump);
      h->launches += 1;
    } else if (!h->hp.large) {
      const size_t sm = mid_smem(p, B > 1);
      auto kern = k_corr_mid<RC>;
      if (int rc2 = allow_smem(h, kern, sm)) return rc2;
      StageTimer timer(h, kStageCorrCols, 1);
      GNSSACQ_LAUNCH(kern, dim3(R * dc), dim3(kThreads), sm, h->stream, p, h->d_X.as<float2>(),
                     h->d_C.as<float2>(), R, B, D, d0, n_lags, scale, h->d_parts.as<Part>(), d_qdump);
      h->launches += 1;
    } else {
      size_t smr = rows_smem(p);
      const size_t smc = cols_smem(p, B > 1);
      const bool gt = p.gt != 0;
      corr_rows_fn kr = spec_on(h) ? find_rows_kernel(p.s2, gt) : nullptr;
      corr_cols_fn kc = spec_on(h) ? find_cols_kernel(p.s1, B > 1) : nullptr;
      int tr = kThreads, tcn = kThreads, row_tile = kTileW;
      if (spec_on(h) && (h->small_ctas & 1)) {
        const RowsSmall rs = find_rows_small(p.s2, gt);
        if (rs.fn) { kr = rs.fn; tr = rs.threads; smr = rs.smem; row_tile = kRowsSmallTile; }
      }
      if (spec_on(h) && (h->small_ctas & 2)) {
        const ColsSmall cs = find_cols_small(p.s1, B > 1);
        if (cs.fn) { kc = cs.fn; tcn = cs.threads; }
      }
      if (!kr) kr = k_corr_rows<RC>;
      if (!kc) kc = k_corr_cols<RC>;
      if (int rc2 = allow_smem(h, kr, smr)) return rc2;
      if (int rc2 = allow_smem(h, kc, smc)) return rc2;
      const int nrt = (p.N1 + row_tile - 1) / row_tile;
      const int units = R * dc;
      const bool two_lanes = h->overlap && units > Uc;
      if (two_lanes) {
        // fork: both lanes wait for everything queued so far on the main stream (the forward FFTs)
        StageTimer timer(h, kStageCorrCols, 0);        // whole correlate section, as seen by the main stream
        CU(cudaEventRecord(h->ev_fork, h->stream));
        for (int l = 0; l < h->nlanes; ++l) CU(cudaStreamWaitEvent(h->lane[l], h->ev_fork, 0));
        int k = 0, nl = 0;
        for (int u0 = 0; u0 < units; u0 += Uc, ++k) {
          const int uc = std::min(Uc, units - u0);
          cudaStream_t st = h->lane[k % h->nlanes];
          float2* scr = h->d_scratch_lane[k % h->nlanes].as<float2>();
          GNSSACQ_LAUNCH(kr, dim3(nrt, B, uc), dim3(tr), smr, st,
                         p, h->d_X.as<float2>(), h->d_C.as<float2>(), R, B, u0, scr);
          GNSSACQ_LAUNCH(kc, dim3(uc, ntiles), dim3(tcn), smc, st, p, scr, R, B, D, d0, u0, n_lags, scale,
                         ntiles, h->d_parts.as<Part>(), d_qdump, h->d_hint.as<unsigned>());
          nl += 2;
        }
        h->launches += nl;
        h->prof_launches[kStageCorrCols] += h->profiling ? nl : 0;
        for (int l = 0; l < h->nlanes; ++l) {          // join
          CU(cudaEventRecord(h->ev_join[l], h->lane[l]));
          CU(cudaStreamWaitEvent(h->stream, h->ev_join[l], 0));
        }
      } else {
        // one stream: the handle's scratch, or — when the search as a whole alternates over lanes and only this
        // (last, short) Doppler chunk does not — the first lane's buffer, which has the same size
        float2* scr1 = h->d_scratch.cap >= (size_t)std::min(Uc, units) * B * p.N * sizeof(float2) ? h->d_scratch.as<float2>()
                                                                                                 : h->d_scratch_lane[0].as<float2>();
        for (int u0 = 0; u0 < units; u0 += Uc) {
          const int uc = std::min(Uc, units - u0);
          {
            StageTimer timer(h, kStageCorrRows, 1);
            GNSSACQ_LAUNCH(kr, dim3(nrt, B, uc), dim3(tr), smr, h->stream,
                           p, h->d_X.as<float2>(), h->d_C.as<float2>(), R, B, u0, scr1);
          }
          StageTimer timer(h, kStageCorrCols, 1);
          GNSSACQ_LAUNCH(kc, dim3(uc, ntiles), dim3(tcn), smc, h->stream, p,
                         scr1, R, B, D, d0, u0, n_lags, scale, ntiles,
                         h->d_parts.as<Part>(), d_qdump, h->d_hint.as<unsigned>());
          h->launches += 2;
        }
      }
    }
    CU(cudaGetLastError());
    return 0;
  });
}

int run_search(gnssacq* h, const double* nco_freq, int D, int stride, int B, int normalize, int n_lags,
               Record* d_out, float* d_qdump, int group_len = 0) {
  if (!h->d_x) return fail(GNSSACQ_ESTATE, "gnssacq_set_signal has not been called");
  if (h->R <= 0) return fail(GNSSACQ_ESTATE, "gnssacq_set_replicas has not been called");
  if (!nco_freq || D <= 0 || B <= 0 || stride < 0) return fail(GNSSACQ_EINVAL, "bad search arguments");
  const int N = h->N, R = h->R;
  const int Nuser = h->embed_n ? h->embed_n : N;              // the caller's transform length (= number of lags)
  if ((int64_t)(B - 1) * stride + Nuser > h->n_x)
    return fail(GNSSACQ_EINVAL, "capture too short: need (n_blocks-1)*block_stride + N = " +
                                    std::to_string((int64_t)(B - 1) * stride + Nuser) + " samples, have " + std::to_string(h->n_x));
  if (n_lags <= 0 || n_lags > Nuser) n_lags = Nuser;
  if (group_len <= 0) group_len = D;
  if (D % group_len != 0 || D / group_len > 65535) return fail(GNSSACQ_EINVAL, "the Doppler list is not a whole number of groups");
  if (B > 65535) return fail(GNSSACQ_EINVAL, "n_blocks too large");
  const DevPlan& p = h->dp;
  const bool large = h->hp.large;
  const V3Setup v3 = large ? v3_setup(h, B > 1, d_qdump != nullptr) : V3Setup();
  // parts per (replica, Doppler): one per tile, or per tile and warp (single-slot columns kernel)
  const int ntiles = v3.on ? v3.ntiles * (v3.f.fn ? 1 : v3.c.parts_per_tile) : (large ? (p.N2 + kTileW - 1) / kTileW : 1);
  const float scale = 1.0f / (float)N;
  const size_t tbytes = (size_t)N * sizeof(float2);

  if (int rc = h->d_freq.ensure((size_t)D * sizeof(double))) return rc;
  CU(cudaMemcpyAsync(h->d_freq.p, nco_freq, (size_t)D * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (int rc = h->d_parts.ensure((size_t)R * D * ntiles * sizeof(Part))) return rc;
  if (large) {
    if (int rc = h->d_hint.ensure((size_t)R * D * sizeof(unsigned))) return rc;
    CU(cudaMemsetAsync(h->d_hint.p, 0, (size_t)R * D * sizeof(unsigned), h->stream));
  }

  int Dc = (int)std::max<size_t>(1, std::min<size_t>((size_t)D, h->xchunk_bytes / (tbytes * B)));
  Dc = std::max(1, std::min(Dc, 65535 / B));
  // units of a Doppler chunk index grids and 32-bit task counters: keep R * Dc (x tiles) well inside int
  if ((long long)R * Dc > (1ll << 24)) Dc = (int)std::max<long long>(1, (1ll << 24) / R);
  if (R > (1 << 24)) return fail(GNSSACQ_EINVAL, "too many replicas for one search");
  if (int rc = h->d_X.ensure((size_t)Dc * B * tbytes)) return rc;
  int Uc = 0;
  if (large && !v3.on) {
    // Units per launch: enough to keep the scratch L2-resident when that still fills the GPU,
    // but never fewer than two waves of columns-kernel CTAs (ntiles x units) — with many
    // non-coherent blocks one unit's scratch alone exceeds L2, and spilling it to HBM
    // (16 B per cell-block) costs far less than an under-filled grid.
    const size_t budget = h->overlap ? h->scratch_bytes / h->nlanes : h->scratch_bytes;
    const size_t unit_bytes = tbytes * B;
    // The 128-thread columns kernel runs 4 CTAs per SM: size a launch to just under one full wave
    // of them — 95 %, so that the other lane's rows kernel finds free slots at once — not just over
    // (measured on config 2, profiles/r02_units_per_chunk_sweep.log: 20 units 2.94 ms, 21-64 units 3.03-3.12 ms).
    const bool small_cols = spec_on(h) && (h->small_ctas & 2) && find_cols_small(p.s1, B > 1).fn != nullptr;
    const size_t fill = small_cols ? std::max<size_t>(1, (size_t)(4 * h->num_sms * 19 / 20) / ntiles)
                                   : (size_t)(4 * h->num_sms + ntiles - 1) / ntiles;
    size_t uc = std::max(budget / unit_bytes, fill);
    uc = std::min(uc, std::max<size_t>(1, ((size_t)3 << 30) / unit_bytes));      // hard cap 3 GiB per lane
    if (h->force_uc > 0) uc = (size_t)h->force_uc;
    Uc = (int)std::max<size_t>(1, std::min<size_t>((size_t)R * Dc, uc));
    Uc = std::min(Uc, 65535);
    if (h->overlap && (long long)R * Dc > Uc) {          // chunks alternate over the lanes' own buffers
      for (int l = 0; l < h->nlanes; ++l)
        if (int rc = h->d_scratch_lane[l].ensure((size_t)Uc * B * tbytes)) return rc;
    } else if (int rc = h->d_scratch.ensure((size_t)Uc * B * tbytes)) return rc;
  }
  for (int d0 = 0; d0 < D; d0 += Dc) {
    const int dc = std::min(Dc, D - d0);
    if (int rc = forward<0>(h, nullptr, h->d_freq.as<double>() + d0, stride, B, dc * B, h->d_X.as<float2>())) return rc;
    if (v3.on && v3.f.fn) {
      if (int rc = correlate_fused(h, v3, B, D, d0, dc, n_lags, scale, d_qdump)) return rc;
    } else if (v3.on) {
      if (int rc = correlate_chunk_v3(h, v3, B, D, d0, dc, n_lags, scale, d_qdump)) return rc;
    } else if (int rc = correlate_chunk(h, B, D, d0, dc, Uc, n_lags, scale, ntiles, d_qdump)) return rc;
  }
  {
    StageTimer timer(h, kStageFinalize, 1);
    GNSSACQ_LAUNCH(k_finalize, dim3(R, D / group_len), dim3(128), 0, h->stream, h->d_parts.as<Part>(), D, group_len, ntiles, Nuser, normalize, d_out);
  }
  h->launches += 1;
  CU(cudaGetLastError());
  return 0;
}

}  // namespace

namespace acq {
#ifndef GNSSACQ_EMU_BUILD
int encode_tensor_map_3d_u64(TensorMap* out, const void* base, const unsigned long long dims[3],
                             const unsigned long long strides_bytes[2], const unsigned box[3]) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static_assert(sizeof(CUtensorMap) == sizeof(TensorMap), "tensor map size");
  static const encode_fn encode = [] {                       // resolved once per process (thread-safe)
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess) fn = nullptr;
    return reinterpret_cast<encode_fn>(fn);
  }();
  if (!encode) return -1;
  const cuuint64_t d[3] = {dims[0], dims[1], dims[2]};
  const cuuint64_t st[2] = {strides_bytes[0], strides_bytes[1]};
  const cuuint32_t bx[3] = {box[0], box[1], box[2]};
  const cuuint32_t es[3] = {1, 1, 1};
  // L2 promotion 128 B even for 64-byte box rows: the other half of the line belongs to the neighbouring
  // column tile, which (tile-major tasks) another CTA fetches shortly after — measured 5 % faster than
  // 64 B promotion although lts__t_bytes is 30 % higher (profiles/README.md r03g)
  const CUresult r = encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(base), d, st, bx, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}
#else
int encode_tensor_map_3d_u64(TensorMap* out, const void* base, const unsigned long long dims[3],
                             const unsigned long long strides_bytes[2], const unsigned box[3]) {
  EmuTensorMap m{};
  m.base = static_cast<const unsigned char*>(base);
  m.rank = 3; m.esize = 8;
  m.stride[0] = 8; m.stride[1] = (long long)strides_bytes[0]; m.stride[2] = (long long)strides_bytes[1];
  for (int i = 0; i < 3; ++i) { m.dim[i] = (long long)dims[i]; m.box[i] = (int)box[i]; }
  memset(out, 0, sizeof(*out));
  memcpy(out, &m, sizeof(m));
  return 0;
}
#endif
}  // namespace acq

// ---------------------------------------------------------------- NCCL, bound at run time
// libgnssacq.so does not link against libnccl: the all-gather of the sharded search resolves the
// few entry points it needs with dlopen/dlsym when gnssacq_nccl_init is first called (inside a
// PyTorch process that is the NCCL PyTorch already loaded), so single-GPU users need no NCCL at all.
namespace {
struct NcclUniqueId { char internal[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
// resolved once per process (thread-safe: function-local static); lib == nullptr if NCCL could not be loaded
const NcclApi& nccl_load(std::string* why) {
  static std::string err;
  static const NcclApi api = [] {
    NcclApi a;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib) break;
    }
    if (!a.lib) { const char* e = dlerror(); err = std::string("NCCL is not available: ") + (e ? e : "dlopen failed"); return a; }
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.lib, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.lib, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.lib, "ncclCommDestroy"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(a.lib, "ncclAllGather"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.lib, "ncclGetErrorString"));
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather) { a = NcclApi(); err = "libnccl lacks an expected entry point"; }
    return a;
  }();
  if (why) *why = err;
  return api;
}
int nccl_api(const NcclApi** out) {
  std::string why;
  const NcclApi& api = nccl_load(&why);
  if (!api.lib) return fail(GNSSACQ_ESTATE, why);
  *out = &api;
  return 0;
}
int nccl_fail(const NcclApi* api, const char* what, int rc) {
  return fail(GNSSACQ_ECUDA, std::string(what) + ": " + (api->GetErrorString ? api->GetErrorString(rc) : "NCCL error") + " (" + std::to_string(rc) + ")");
}
}  // namespace

// =============================================================================== C ABI
extern "C" {

int gnssacq_nccl_unique_id(void* id128) {
  if (!id128) return fail(GNSSACQ_EINVAL, "NULL argument");
  const NcclApi* api;
  if (int rc = nccl_api(&api)) return rc;
  NcclUniqueId id;
  if (int rc = api->GetUniqueId(&id)) return nccl_fail(api, "ncclGetUniqueId", rc);
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int gnssacq_nccl_init(gnssacq_t* h, const void* id128, int32_t rank, int32_t world) {
  if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return fail(GNSSACQ_EINVAL, "bad communicator arguments");
  const NcclApi* api;
  if (int rc = nccl_api(&api)) return rc;
  CU(cudaSetDevice(h->device));
  if (h->nccl_comm) { api->CommDestroy(h->nccl_comm); h->nccl_comm = nullptr; }
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  if (int rc = api->CommInitRank(&h->nccl_comm, world, id, rank)) { h->nccl_comm = nullptr; return nccl_fail(api, "ncclCommInitRank", rc); }
  h->nccl_rank = rank;
  h->nccl_world = world;
  return 0;
}

int gnssacq_search_sharded(gnssacq_t* h, const double* nco_freq, int32_t D, int32_t block_stride, int32_t n_blocks,
                           int32_t normalize, int32_t n_lags, float* metric, int32_t* lag, int32_t* dbin) {
  if (!h || !nco_freq || !metric || !lag || !dbin || D <= 0) return fail(GNSSACQ_EINVAL, "bad search arguments");
  CU(cudaSetDevice(h->device));
  if (h->R <= 0) return fail(GNSSACQ_ESTATE, "gnssacq_set_replicas has not been called");
  const int world = h->nccl_comm ? h->nccl_world : 1, rank = h->nccl_comm ? h->nccl_rank : 0;
  const int R = h->R;
  const int base = D / world, extra = D % world;
  const int lo = rank * base + std::min(rank, extra), cnt = base + (rank < extra ? 1 : 0);
  if (int rc = h->d_rec.ensure((size_t)R * sizeof(Record))) return rc;
  if (int rc = h->d_allrec.ensure((size_t)world * R * sizeof(Record))) return rc;
  if (int rc = h->d_merged.ensure((size_t)R * sizeof(Record))) return rc;
  if (cnt > 0) {
    if (int rc = run_search(h, nco_freq + lo, cnt, block_stride, n_blocks, normalize, n_lags, h->d_rec.as<Record>(), nullptr)) return rc;
  } else {
    std::vector<Record> none(R);
    for (auto& r : none) { r.metric = 0.f; r.lag = 0; r.dbin = -1; r.pad = 0; }
    CU(cudaMemcpyAsync(h->d_rec.p, none.data(), none.size() * sizeof(Record), cudaMemcpyHostToDevice, h->stream));
    SYNC_MAIN(h);
  }
  if (world > 1) {
    const NcclApi* api;
    if (int rc = nccl_api(&api)) return rc;
    if (int rc = api->AllGather(h->d_rec.p, h->d_allrec.p, (size_t)R * 4, /*ncclInt32*/ 2, h->nccl_comm, h->stream))
      return nccl_fail(api, "ncclAllGather", rc);
  } else {
    CU(cudaMemcpyAsync(h->d_allrec.p, h->d_rec.p, (size_t)R * sizeof(Record), cudaMemcpyDeviceToDevice, h->stream));
  }
  GNSSACQ_LAUNCH(k_merge_records, dim3((R + 127) / 128), dim3(128), 0, h->stream, h->d_allrec.as<Record>(), world, R, D, h->d_merged.as<Record>());
  h->launches += 1;
  CU(cudaGetLastError());
  std::vector<Record> rec(R);
  CU(cudaMemcpyAsync(rec.data(), h->d_merged.p, rec.size() * sizeof(Record), cudaMemcpyDeviceToHost, h->stream));
  SYNC_MAIN(h);
  for (int r = 0; r < R; ++r) { metric[r] = rec[r].metric; lag[r] = rec[r].lag; dbin[r] = rec[r].dbin; }
  return 0;
}


const char* gnssacq_last_error(void) { return g_err.c_str(); }

int gnssacq_create(int device, gnssacq_t** out) {
  if (!out) return fail(GNSSACQ_EINVAL, "out is NULL");
  *out = nullptr;
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(GNSSACQ_EINVAL, "no such CUDA device " + std::to_string(device));
  CU(cudaSetDevice(device));
  gnssacq* h = new (std::nothrow) gnssacq();
  if (!h) return fail(GNSSACQ_ENOMEM, "out of host memory");
  h->device = device;
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  for (int l = 0; l < gnssacq::kMaxLanes && e == cudaSuccess; ++l) {
    e = cudaStreamCreateWithFlags(&h->lane[l], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join[l], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_x, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_before_x, cudaEventDisableTiming);
  if (e != cudaSuccess) { delete h; return fail(GNSSACQ_ECUDA, cudaGetErrorString(e)); }
  h->smem_optin = prop.sharedMemPerBlockOptin;
  h->num_sms = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
  h->stream = h->own_stream;
  h->nco_c128.resize(2 * kNcoSize);
  for (int k = 0; k < kNcoSize; ++k) {
    const double th = (2.0 * M_PI * (double)k) * (1.0 / kNcoSize);
    h->nco_c128[2 * k] = cos(th);
    h->nco_c128[2 * k + 1] = sin(th);
  }
  if (int rc = upload_nco(h)) { gnssacq_destroy(h); return rc; }
  *out = h;
  return 0;
}

int gnssacq_destroy(gnssacq_t* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  join_capture_copy(h);
  cudaStreamSynchronize(h->stream);
  for (DevBuf* b : {&h->d_nco_f32, &h->d_nco_f64, &h->d_x_own, &h->d_tw1, &h->d_tw2, &h->d_twm, &h->d_twm_inv, &h->d_maps, &h->d_cube0, &h->d_cube1, &h->d_C, &h->d_X,
                    &h->d_scratch, &h->d_parts, &h->d_freq, &h->d_rec, &h->d_q, &h->d_tmp, &h->d_raw, &h->d_ext,
                    &h->d_y1, &h->d_z, &h->d_fir, &h->d_pre128, &h->d_chips, &h->d_base, &h->d_bank, &h->d_v3tab, &h->d_hint, &h->d_repext, &h->d_allrec, &h->d_merged, &h->d_fsync})
    b->release();
  for (auto& sp : h->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  for (auto e : h->event_pool) cudaEventDestroy(e);
  for (int l = 0; l < gnssacq::kMaxLanes; ++l) {
    if (h->lane[l]) { cudaStreamSynchronize(h->lane[l]); cudaStreamDestroy(h->lane[l]); }
    if (h->ev_join[l]) cudaEventDestroy(h->ev_join[l]);
    h->d_scratch_lane[l].release();
  }
  if (h->nccl_comm) { const NcclApi& api = nccl_load(nullptr); if (api.CommDestroy) api.CommDestroy(h->nccl_comm); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  if (h->ev_x) cudaEventDestroy(h->ev_x);
  if (h->ev_before_x) cudaEventDestroy(h->ev_before_x);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return 0;
}

int gnssacq_set_stream(gnssacq_t* h, void* cuda_stream) {
  if (!h) return fail(GNSSACQ_EINVAL, "handle is NULL");
  SYNC_MAIN(h);
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return 0;
}

int gnssacq_set_nco_table(gnssacq_t* h, const double* table_c128) {
  if (!h || !table_c128) return fail(GNSSACQ_EINVAL, "NULL argument");
  CU(cudaSetDevice(h->device));
  h->nco_c128.assign(table_c128, table_c128 + 2 * kNcoSize);
  return upload_nco(h);
}

int gnssacq_set_signal(gnssacq_t* h, const float* iq, int64_t n) {
  if (!h || !iq || n <= 0) return fail(GNSSACQ_EINVAL, "bad signal arguments");
  CU(cudaSetDevice(h->device));
  if (int rc = h->d_x_own.ensure((size_t)n * sizeof(float2))) return rc;
  // on the copy stream, behind everything queued so far (earlier searches still read the buffer)
  CU(cudaEventRecord(h->ev_before_x, h->stream));
  CU(cudaStreamWaitEvent(h->copy_stream, h->ev_before_x, 0));
  CU(cudaMemcpyAsync(h->d_x_own.p, iq, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, h->copy_stream));
  CU(cudaEventRecord(h->ev_x, h->copy_stream));
  h->x_pending = true;
  h->d_x = h->d_x_own.as<float2>();
  h->n_x = n;
  return 0;
}

int gnssacq_set_signal_device(gnssacq_t* h, const void* dev_iq, int64_t n) {
  if (!h || !dev_iq || n <= 0) return fail(GNSSACQ_EINVAL, "bad signal arguments");
  h->d_x = static_cast<const float2*>(dev_iq);
  h->n_x = n;
  return 0;
}

static int replicas_from_device(gnssacq_t* h, const float* d_rep, int32_t R, int32_t N) {
  // Lengths the planner cannot factor (a prime factor outside {2,3,5,7,11,13,31}, or no split into
  // two factors <= 1024) run embedded in the next power of two M >= 2N-1: periodically extended
  // replica, zero-padded capture blocks, lags 0..N-1 kept — the same circular correlation, at the
  // price of a longer transform and the generic kernels.
  {
    int want = 0;
    if (!plannable(N)) {
      long long M = 1;
      while (M < 2ll * N - 1) M <<= 1;
      if (N < 4 || M > (long long)kMaxSub * kMaxSub)
        return fail(GNSSACQ_EINVAL, "FFT length " + std::to_string(N) + " is neither plannable (prime factors 2, 3, 5, 7, 11, 13, 31, two factors <= 1024) "
                                    "nor short enough to embed in a power-of-two transform (N <= 524288)");
      want = N;
      N = (int32_t)M;
    }
    if (want != h->embed_n) { h->embed_n = want; h->plan_dirty = true; }
  }
  if (int rc = upload_plan(h, N)) return rc;
  if (h->embed_n) {
    if (int rc = h->d_repext.ensure((size_t)R * N * sizeof(float))) return rc;
    GNSSACQ_LAUNCH(k_extend_replicas, dim3(std::min((N + kThreads - 1) / kThreads, 1024), R), dim3(kThreads), 0, h->stream, d_rep, h->embed_n, N,
                   h->d_repext.as<float>());
    h->launches += 1;
    CU(cudaGetLastError());
    d_rep = h->d_repext.as<float>();
  }
  if (int rc = h->d_C.ensure((size_t)R * N * sizeof(float2))) return rc;
  h->R = 0; h->N = N;
  // grid.y of the large path is limited to 65535 transforms per launch
  for (int r0 = 0; r0 < R; r0 += 32768) {
    const int rc_n = std::min(32768, R - r0);
    if (int rc = forward<1>(h, d_rep + (size_t)r0 * N, nullptr, 0, 1, rc_n, h->d_C.as<float2>() + (size_t)r0 * N)) return rc;
  }
  h->R = R;
  return 0;
}

int gnssacq_set_replicas(gnssacq_t* h, const float* replicas, int32_t R, int32_t N) {
  if (!h || !replicas || R <= 0 || N <= 0) return fail(GNSSACQ_EINVAL, "bad replica arguments");
  CU(cudaSetDevice(h->device));
  const size_t nel = (size_t)R * N;
  if (int rc = h->d_tmp.ensure(nel * sizeof(float))) return rc;
  CU(cudaMemcpyAsync(h->d_tmp.p, replicas, nel * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  return replicas_from_device(h, h->d_tmp.as<float>(), R, N);
}

static int replicas_from_device_i8(gnssacq_t* h, const signed char* d_i8, int32_t R, int32_t N) {
  const size_t nel = (size_t)R * N;
  if (int rc = h->d_tmp.ensure(nel * sizeof(float))) return rc;
  const int blocks = (int)std::min<size_t>((nel + kThreads - 1) / kThreads, (size_t)h->num_sms * 16);
  GNSSACQ_LAUNCH(k_i8_to_f32, dim3(blocks), dim3(kThreads), 0, h->stream, d_i8, (long long)nel, h->d_tmp.as<float>());
  h->launches += 1;
  CU(cudaGetLastError());
  return replicas_from_device(h, h->d_tmp.as<float>(), R, N);
}

int gnssacq_set_replicas_i8(gnssacq_t* h, const int8_t* replicas, int32_t R, int32_t N) {
  if (!h || !replicas || R <= 0 || N <= 0) return fail(GNSSACQ_EINVAL, "bad replica arguments");
  CU(cudaSetDevice(h->device));
  const size_t nel = (size_t)R * N;
  if (int rc = h->d_raw.ensure(nel)) return rc;
  CU(cudaMemcpyAsync(h->d_raw.p, replicas, nel, cudaMemcpyHostToDevice, h->stream));
  return replicas_from_device_i8(h, h->d_raw.as<signed char>(), R, N);
}

int gnssacq_set_replicas_i8_device(gnssacq_t* h, const void* device_replicas_i8, int32_t R, int32_t N) {
  if (!h || !device_replicas_i8 || R <= 0 || N <= 0) return fail(GNSSACQ_EINVAL, "bad replica arguments");
  CU(cudaSetDevice(h->device));
  return replicas_from_device_i8(h, static_cast<const signed char*>(device_replicas_i8), R, N);
}

int gnssacq_set_replicas_device(gnssacq_t* h, const void* device_replicas, int32_t R, int32_t N) {
  if (!h || !device_replicas || R <= 0 || N <= 0) return fail(GNSSACQ_EINVAL, "bad replica arguments");
  CU(cudaSetDevice(h->device));
  return replicas_from_device(h, static_cast<const float*>(device_replicas), R, N);
}

int gnssacq_set_replicas_from_chips(gnssacq_t* h, const int8_t* chips01, int32_t R, int32_t L, int32_t n, int32_t N,
                                    double base, double incr, int32_t boc, double base2) {
  if (!h || !chips01 || R <= 0 || L <= 0 || n <= 0 || N < n) return fail(GNSSACQ_EINVAL, "bad replica arguments");
  if (R > 65535) return fail(GNSSACQ_EINVAL, "too many replicas for one call");
  CU(cudaSetDevice(h->device));
  if (int rc = h->d_chips.ensure((size_t)R * L)) return rc;
  if (int rc = h->d_tmp.ensure((size_t)R * N * sizeof(float))) return rc;
  CU(cudaMemcpyAsync(h->d_chips.p, chips01, (size_t)R * L, cudaMemcpyHostToDevice, h->stream));
  GNSSACQ_LAUNCH(k_build_replicas, dim3((N + kThreads - 1) / kThreads, R), dim3(kThreads), 0, h->stream,
                 h->d_chips.as<signed char>(), L, n, N, base, incr, boc, base2, h->d_tmp.as<float>());
  h->launches += 1;
  CU(cudaGetLastError());
  return replicas_from_device(h, h->d_tmp.as<float>(), R, N);
}

int gnssacq_correlate_bank(gnssacq_t* h, const int8_t* chips01, int32_t L, double nco_freq, int32_t n,
                           int32_t n_blocks, int32_t block_stride, const double* base, int32_t H, double incr,
                           double* out_c128) {
  if (!h || !chips01 || !base || !out_c128 || L <= 0 || n <= 0 || n_blocks <= 0 || block_stride < 0 || H <= 0)
    return fail(GNSSACQ_EINVAL, "bad correlator-bank arguments");
  if (n_blocks > 65535) return fail(GNSSACQ_EINVAL, "n_blocks too large");
  if (!h->d_x) return fail(GNSSACQ_ESTATE, "gnssacq_set_signal has not been called");
  if (int rc = join_capture_copy(h)) return rc;
  if ((int64_t)(n_blocks - 1) * block_stride + n > h->n_x)
    return fail(GNSSACQ_EINVAL, "capture too short: need (n_blocks-1)*block_stride + n = " +
                                    std::to_string((int64_t)(n_blocks - 1) * block_stride + n) + " samples, have " + std::to_string(h->n_x));
  CU(cudaSetDevice(h->device));
  const size_t nhb = (size_t)H * n_blocks;
  if (int rc = h->d_chips.ensure((size_t)L)) return rc;
  if (int rc = h->d_base.ensure(nhb * sizeof(double))) return rc;
  if (int rc = h->d_bank.ensure(nhb * sizeof(double2))) return rc;
  if (int rc = h->d_scratch.ensure((size_t)n_blocks * n * sizeof(float2))) return rc;
  CU(cudaMemcpyAsync(h->d_chips.p, chips01, (size_t)L, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->d_base.p, base, nhb * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  GNSSACQ_LAUNCH(k_bank_wipeoff, dim3((n + kThreads - 1) / kThreads, n_blocks), dim3(kThreads), 0, h->stream,
                 h->d_x, h->d_nco_f32.as<float2>(), nco_freq, n, block_stride, h->d_scratch.as<float2>());
  h->launches += 1;
  // grid.x carries the hypotheses in slices (any H), grid.y the blocks
  for (int h0 = 0; h0 < H; h0 += 32768) {
    const int hc = std::min(32768, H - h0);
    GNSSACQ_LAUNCH(k_corr_bank, dim3(hc, n_blocks), dim3(kThreads), 0, h->stream, h->d_scratch.as<float2>(),
                   h->d_chips.as<signed char>(), L, n, n_blocks, h->d_base.as<double>() + (size_t)h0 * n_blocks, incr,
                   h->d_bank.as<double2>() + (size_t)h0 * n_blocks);
    h->launches += 1;
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out_c128, h->d_bank.p, nhb * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
  SYNC_MAIN(h);
  return 0;
}

int gnssacq_correlate_epl(gnssacq_t* h, const float* x_c64, int32_t nx, int32_t n, const int8_t* chips01, int32_t ncodes, int32_t L,
                          int32_t mode, const double* params, int32_t H, const int32_t* xsel, const int32_t* csel,
                          const double* start, const double* incr, double* out_c128) {
  if (!h || !x_c64 || !chips01 || !xsel || !csel || !start || !incr || !out_c128 || nx <= 0 || n <= 0 || ncodes <= 0 || L <= 0 || H <= 0)
    return fail(GNSSACQ_EINVAL, "bad correlator arguments");
  if (mode < 0 || mode > 3 || (mode >= 1 && !params)) return fail(GNSSACQ_EINVAL, "mode must be 0..3 (modes 1-3 need params)");
  for (int i = 0; i < H; ++i)
    if (xsel[i] < 0 || xsel[i] >= nx || csel[i] < 0 || csel[i] >= ncodes || !(incr[i] > 0.0)) return fail(GNSSACQ_EINVAL, "hypothesis out of range");
  CU(cudaSetDevice(h->device));
  const size_t xb = (size_t)nx * n * sizeof(float2), cb = (size_t)ncodes * L, pb = 37 * sizeof(double);
  const size_t hb_i = (size_t)H * sizeof(int), hb_d = (size_t)H * sizeof(double);
  if (int rc = h->d_tmp.ensure(xb)) return rc;
  if (int rc = h->d_chips.ensure(cb)) return rc;
  if (int rc = h->d_base.ensure(pb + 2 * hb_d + 2 * hb_i)) return rc;
  if (int rc = h->d_bank.ensure((size_t)H * sizeof(double2))) return rc;
  unsigned char* base = h->d_base.as<unsigned char>();
  double* d_params = reinterpret_cast<double*>(base);
  double* d_start = d_params + 37;
  double* d_incr = d_start + H;
  int* d_xsel = reinterpret_cast<int*>(d_incr + H);
  int* d_csel = d_xsel + H;
  double hp[37] = {1.0, 1.0, 0.0, 0.0};
  if (params) memcpy(hp, params, (mode == 3 ? 37 : 4) * sizeof(double));
  CU(cudaMemcpyAsync(h->d_tmp.p, x_c64, xb, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->d_chips.p, chips01, cb, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(d_params, hp, pb, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(d_start, start, hb_d, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(d_incr, incr, hb_d, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(d_xsel, xsel, hb_i, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(d_csel, csel, hb_i, cudaMemcpyHostToDevice, h->stream));
  SYNC_MAIN(h);                    // hp is a local
  GNSSACQ_LAUNCH(k_correlate_epl, dim3(H), dim3(kThreads), 0, h->stream, h->d_tmp.as<float2>(), n, h->d_chips.as<signed char>(), L, mode,
                 d_params, d_xsel, d_csel, d_start, d_incr, h->d_bank.as<double2>());
  h->launches += 1;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out_c128, h->d_bank.p, (size_t)H * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
  SYNC_MAIN(h);
  return 0;
}

int gnssacq_set_option(gnssacq_t* h, const char* name, int32_t value) {
  if (!h || !name) return fail(GNSSACQ_EINVAL, "NULL argument");
  if (std::string(name) == "specialized_kernels") {     // changes the spectrum layout: replicas must be set again
    if (h->use_spec != (value != 0)) { h->R = 0; h->plan_dirty = true; }
    h->use_spec = value != 0;
    return 0;
  }
  if (std::string(name) == "gt_split") {                // changes the spectrum layout: replicas must be set again
    if (h->use_gt != (value != 0)) { h->R = 0; h->plan_dirty = true; }
    h->use_gt = value != 0;
    return 0;
  }
  if (std::string(name) == "v3") { h->use_v3 = value != 0; return 0; }
  if (std::string(name) == "fused") { h->use_fused = value != 0; return 0; }
  if (std::string(name) == "fused_rc") { h->fused_rc = value; return 0; }
  if (std::string(name) == "fused_g") { h->fused_g = value; return 0; }
  if (std::string(name) == "fused_sets") { h->fused_sets = value; return 0; }
  if (std::string(name) == "fused_ctas") { h->fused_ctas = value; return 0; }
  if (std::string(name) == "fused_tpt") { h->fused_tpt = value; return 0; }
  if (std::string(name) == "v3_rows") { h->v3_rows_variant = value; return 0; }
  if (std::string(name) == "v3_cols") { h->v3_cols_variant = value; return 0; }
  if (std::string(name) == "fwd_v6") { h->fwd_v6 = value != 0; return 0; }
  if (std::string(name) == "v3_rc") { h->v3_rc = value; return 0; }
  if (std::string(name) == "v3_g") { h->v3_g = value; return 0; }
  if (std::string(name) == "overlap_chunks") { h->overlap = value != 0; return 0; }
  if (std::string(name) == "small_ctas") { h->small_ctas = value; return 0; }
  if (std::string(name) == "units_per_chunk") { h->force_uc = value; return 0; }
  if (std::string(name) == "split_n1") { h->force_n1 = value; h->R = 0; h->plan_dirty = true; return 0; }   // replicas must be set again
  if (std::string(name) == "disable_radix") {          // value 0 clears the list
    if (value == 0) h->disabled_radices = 0;
    else if (value > 0 && value < 64) h->disabled_radices |= 1ull << value;
    else return fail(GNSSACQ_EINVAL, "disable_radix: 0..63");
    h->R = 0; h->plan_dirty = true;
    return 0;
  }
  if (std::string(name) == "lanes") {
    if (value < 1 || value > gnssacq::kMaxLanes) return fail(GNSSACQ_EINVAL, "lanes must be 1..4");
    h->nlanes = value; h->overlap = value > 1; return 0;
  }
  if (std::string(name) == "scratch_mb") { if (value < 1) return fail(GNSSACQ_EINVAL, "scratch_mb"); h->scratch_bytes = (size_t)value << 20; return 0; }
  if (std::string(name) == "xchunk_mb") { if (value < 1) return fail(GNSSACQ_EINVAL, "xchunk_mb"); h->xchunk_bytes = (size_t)value << 20; return 0; }
  return fail(GNSSACQ_EINVAL, std::string("unknown option ") + name);
}

int gnssacq_set_schedule(gnssacq_t* h, int32_t which, const int32_t* radices, int32_t n) {
  if (!h || which < 1 || which > 2 || n < 0 || (n > 0 && !radices)) return fail(GNSSACQ_EINVAL, "bad schedule arguments");
  h->forced_sched[which - 1].assign(radices, radices + n);
  h->R = 0; h->plan_dirty = true;
  return 0;
}

int gnssacq_set_profiling(gnssacq_t* h, int32_t on) {
  if (!h) return fail(GNSSACQ_EINVAL, "handle is NULL");
  h->profiling = on != 0;
  return 0;
}

int gnssacq_get_stage_times(gnssacq_t* h, double* ms4, int64_t* launches4, int32_t reset) {
  if (!h) return fail(GNSSACQ_EINVAL, "handle is NULL");
  CU(cudaSetDevice(h->device));
  SYNC_MAIN(h);
  for (auto& sp : h->spans) {
    float ms = 0.f;
#ifndef GNSSACQ_EMU_BUILD
    cudaEventElapsedTime(&ms, sp.a, sp.b);
#endif
    h->prof_ms[sp.stage] += ms;
    h->event_pool.push_back(sp.a);
    h->event_pool.push_back(sp.b);
  }
  h->spans.clear();
  for (int i = 0; i < 4; ++i) {
    if (ms4) ms4[i] = h->prof_ms[i];
    if (launches4) launches4[i] = h->prof_launches[i];
    if (reset) { h->prof_ms[i] = 0; h->prof_launches[i] = 0; }
  }
  return 0;
}

int gnssacq_search_device(gnssacq_t* h, const double* nco_freq, int32_t D, int32_t block_stride, int32_t n_blocks,
                          int32_t normalize, int32_t n_lags, void* device_records) {
  if (!h || !device_records) return fail(GNSSACQ_EINVAL, "NULL argument");
  CU(cudaSetDevice(h->device));
  return run_search(h, nco_freq, D, block_stride, n_blocks, normalize, n_lags, static_cast<Record*>(device_records), nullptr);
}

int gnssacq_search(gnssacq_t* h, const double* nco_freq, int32_t D, int32_t block_stride, int32_t n_blocks,
                   int32_t normalize, int32_t n_lags, float* metric, int32_t* lag, int32_t* dbin, float* q_dump) {
  if (!h || !metric || !lag || !dbin) return fail(GNSSACQ_EINVAL, "NULL argument");
  CU(cudaSetDevice(h->device));
  if (h->R <= 0) return fail(GNSSACQ_ESTATE, "gnssacq_set_replicas has not been called");
  if (int rc = h->d_rec.ensure((size_t)h->R * sizeof(Record))) return rc;
  float* d_q = nullptr;
  const size_t qn = (size_t)h->R * (size_t)std::max(D, 0) * h->N;
  if (q_dump) {
    if (int rc = h->d_q.ensure(qn * sizeof(float))) return rc;
    d_q = h->d_q.as<float>();
  }
  if (int rc = run_search(h, nco_freq, D, block_stride, n_blocks, normalize, n_lags, h->d_rec.as<Record>(), d_q)) return rc;
  std::vector<Record> rec(h->R);
  CU(cudaMemcpyAsync(rec.data(), h->d_rec.p, rec.size() * sizeof(Record), cudaMemcpyDeviceToHost, h->stream));
  if (q_dump) {
    if (h->embed_n)            // the device grid has the planned length per row; the caller gets lags 0..N-1
      CU(cudaMemcpy2DAsync(q_dump, (size_t)h->embed_n * sizeof(float), d_q, (size_t)h->N * sizeof(float), (size_t)h->embed_n * sizeof(float),
                           (size_t)h->R * D, cudaMemcpyDeviceToHost, h->stream));
    else
      CU(cudaMemcpyAsync(q_dump, d_q, qn * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  }
  SYNC_MAIN(h);
  for (int r = 0; r < h->R; ++r) { metric[r] = rec[r].metric; lag[r] = rec[r].lag; dbin[r] = rec[r].dbin; }
  return 0;
}

int gnssacq_search_grouped(gnssacq_t* h, const double* nco_freq, int32_t D, int32_t group_len, int32_t block_stride,
                           int32_t n_blocks, int32_t normalize, int32_t n_lags, float* metric, int32_t* lag, int32_t* dbin) {
  if (!h || !metric || !lag || !dbin) return fail(GNSSACQ_EINVAL, "NULL argument");
  if (group_len <= 0 || D <= 0 || D % group_len != 0) return fail(GNSSACQ_EINVAL, "the Doppler list is not a whole number of groups");
  CU(cudaSetDevice(h->device));
  if (h->R <= 0) return fail(GNSSACQ_ESTATE, "gnssacq_set_replicas has not been called");
  const size_t nrec = (size_t)h->R * (D / group_len);
  if (int rc = h->d_rec.ensure(nrec * sizeof(Record))) return rc;
  if (int rc = run_search(h, nco_freq, D, block_stride, n_blocks, normalize, n_lags, h->d_rec.as<Record>(), nullptr, group_len)) return rc;
  std::vector<Record> rec(nrec);
  CU(cudaMemcpyAsync(rec.data(), h->d_rec.p, nrec * sizeof(Record), cudaMemcpyDeviceToHost, h->stream));
  SYNC_MAIN(h);
  for (size_t i = 0; i < nrec; ++i) { metric[i] = rec[i].metric; lag[i] = rec[i].lag; dbin[i] = rec[i].dbin; }
  return 0;
}

int gnssacq_mix(gnssacq_t* h, float* iq, int64_t n, double f, double p) {
  if (!h || !iq || n < 0) return fail(GNSSACQ_EINVAL, "bad mix arguments");
  if (n == 0) return 0;
  CU(cudaSetDevice(h->device));
  // dp = int(floor(p*NT*(1<<50))), df likewise (gnsstools/nco.py:33-34)
  const double scale = (double)kNcoSize * (double)(1ll << 50);
  const long long dp0 = (long long)floor(p * scale);
  const long long df = (long long)floor(f * scale);
  if (int rc = h->d_tmp.ensure((size_t)n * sizeof(float2))) return rc;
  CU(cudaMemcpyAsync(h->d_tmp.p, iq, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
  const int blocks = (int)std::min<int64_t>((n + kThreads - 1) / kThreads, (int64_t)h->num_sms * 16);
  GNSSACQ_LAUNCH(k_mix, dim3(blocks), dim3(kThreads), 0, h->stream, h->d_tmp.as<float2>(), (long long)n,
                 (unsigned long long)dp0, (unsigned long long)df, h->d_nco_f64.as<double2>());
  h->launches += 1;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(iq, h->d_tmp.p, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, h->stream));
  SYNC_MAIN(h);
  return 0;
}

int gnssacq_preprocess(gnssacq_t* h, const int8_t* iq, int64_t n, double mix_f, double mix_p,
                       const double* fir, int32_t ntaps, double step, int64_t n_out, double* out_c128) {
  if (!h || !iq || !fir || n <= 0 || ntaps <= 0 || n_out <= 0) return fail(GNSSACQ_EINVAL, "bad preprocess arguments");
  const int edge = 3 * ntaps;                              // filtfilt default padlen = 3*max(len(a),len(b))
  if (n <= edge) return fail(GNSSACQ_EINVAL, "recording shorter than the filter padding (filtfilt would raise)");
  CU(cudaSetDevice(h->device));
  if (int rc = join_capture_copy(h)) return rc;                // this call rewrites the capture buffer
  const long long L = n + 2ll * edge;
  const double scale = (double)kNcoSize * (double)(1ll << 50);
  const long long dp0 = (long long)floor(mix_p * scale);
  const long long df = (long long)floor(mix_f * scale);
  if (int rc = h->d_raw.ensure((size_t)2 * n)) return rc;
  if (int rc = h->d_ext.ensure((size_t)L * sizeof(float2))) return rc;
  if (int rc = h->d_y1.ensure((size_t)L * sizeof(double2))) return rc;
  if (int rc = h->d_z.ensure((size_t)L * sizeof(double2))) return rc;
  if (int rc = h->d_fir.ensure((size_t)ntaps * sizeof(double))) return rc;
  if (int rc = h->d_x_own.ensure((size_t)n_out * sizeof(float2))) return rc;
  if (out_c128) if (int rc = h->d_pre128.ensure((size_t)n_out * sizeof(double2))) return rc;
  CU(cudaMemcpyAsync(h->d_raw.p, iq, (size_t)2 * n, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->d_fir.p, fir, (size_t)ntaps * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  const int blocks = (int)std::min<long long>((L + kThreads - 1) / kThreads, (long long)h->num_sms * 16);
  const double2* tab = h->d_nco_f64.as<double2>();
  GNSSACQ_LAUNCH(k_pre_extend, dim3(blocks), dim3(kThreads), 0, h->stream, h->d_raw.as<signed char>(), (long long)n, edge,
                 (unsigned long long)dp0, (unsigned long long)df, tab, h->d_ext.as<float2>());
  // forward pass where all taps are inside ext; backward pass over what the slice [edge, edge+n) needs
  auto kf = k_pre_fir<1, float2>;
  auto kb = k_pre_fir<-1, double2>;
  GNSSACQ_LAUNCH(kf, dim3(blocks), dim3(kThreads), 0, h->stream, h->d_ext.as<float2>(), h->d_fir.as<double>(), ntaps,
                 (long long)(ntaps - 1), L, h->d_y1.as<double2>());
  GNSSACQ_LAUNCH(kb, dim3(blocks), dim3(kThreads), 0, h->stream, h->d_y1.as<double2>(), h->d_fir.as<double>(), ntaps,
                 (long long)edge, (long long)edge + n, h->d_z.as<double2>());
  const int blocks_o = (int)std::min<long long>((n_out + kThreads - 1) / kThreads, (long long)h->num_sms * 16);
  GNSSACQ_LAUNCH(k_pre_interp, dim3(blocks_o), dim3(kThreads), 0, h->stream, h->d_z.as<double2>(), (long long)n, edge, step,
                 (long long)n_out, h->d_x_own.as<float2>(), out_c128 ? h->d_pre128.as<double2>() : nullptr);
  h->launches += 4;
  CU(cudaGetLastError());
  h->d_x = h->d_x_own.as<float2>();
  h->n_x = n_out;
  if (out_c128) {
    CU(cudaMemcpyAsync(out_c128, h->d_pre128.p, (size_t)n_out * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
    SYNC_MAIN(h);
  }
  return 0;
}

int gnssacq_plan_info(gnssacq_t* h, int32_t* N, int32_t* N1, int32_t* N2, int32_t* large) {
  if (!h || h->hp.N == 0) return fail(GNSSACQ_ESTATE, "no plan yet");
  if (N) *N = h->hp.N;
  if (N1) *N1 = h->hp.N1;
  if (N2) *N2 = h->hp.N2;
  if (large) *large = h->hp.large ? 1 : 0;
  return 0;
}

int64_t gnssacq_launch_count(gnssacq_t* h) { return h ? h->launches : 0; }

int gnssacq_kernel_variant(gnssacq_t* h) {
  if (!h || h->hp.N == 0) return fail(GNSSACQ_ESTATE, "no plan yet");
  if (h->embed_n) return 128;                                   // embedded length: generic kernels on a power-of-two plan
  if (h->hp.cube && h->use_spec) return 4;
  if (!h->hp.large || !h->use_spec) return 0;
  return (find_rows_kernel(h->dp.s2, h->hp.gt) ? 1 : 0) | (find_cols_kernel(h->dp.s1, false) ? 2 : 0) | (h->hp.s1.pfa ? 8 : 0) | (h->hp.s2.pfa ? 16 : 0) |
         (h->hp.gt ? 32 : 0) | (v3_setup(h, false, false).on ? 64 : 0) | (v3_setup(h, false, false).f.fn ? 256 : 0);
}

int gnssacq_synchronize(gnssacq_t* h) {
  if (!h) return fail(GNSSACQ_EINVAL, "handle is NULL");
  CU(cudaSetDevice(h->device));
  SYNC_MAIN(h);
  return 0;
}

}  // extern "C"
