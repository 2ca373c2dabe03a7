// A/B harness for the correlate kernels of one plan (default 163680 = 372 x 440): runs the
// one-tile-per-CTA kernels and the pipelined ones on the same random spectra, checks that the
// per-tile results agree bit for bit, and times each with CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo
//        -DSCOLS=S372 -DSROWS=S440 -DNFFT=163680 -o corr_pipe tools/microbench/corr_pipe.cu
#include "../../gnss-dsp-tools_b200/csrc/fft_plan.h"
#include "kernels_pipe.cuh"
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
using namespace acq;

#ifndef SCOLS
#define SCOLS S372
#define SROWS S440
#define NFFT 163680
#endif
#ifndef RTHREADS
#define RTHREADS 256
#endif
#ifndef TTHREADS
#define TTHREADS 128
#define TMINCTAS 7
#endif
#ifndef CTHREADS
#define CTHREADS 128
#define CMINCTAS 4
#define CSPLIT false
#endif
#ifndef BLOCKS
#define BLOCKS 1
#endif
constexpr bool kMulti = BLOCKS > 1;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

static void fill_sub(const HostSubPlan& hs, const float2* tw, SubPlan& sp) {
  sp.F = hs.F; sp.ns = (int)hs.radix.size();
  for (int j = 0; j < kMaxStages; ++j) {
    sp.radix[j] = j < sp.ns ? hs.radix[j] : 1; sp.m[j] = j < sp.ns ? hs.m[j] : 1; sp.tws_off[j] = j < sp.ns ? hs.tws_off[j] : 0;
  }
  sp.tws0_t_off = hs.tws0_t_off; sp.tw = tw; sp.pfa = hs.pfa ? 1 : 0;
}
template <class T> T* dev(const std::vector<T>& v) {
  T* p; CK(cudaMalloc(&p, v.size() * sizeof(T))); CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice)); return p;
}

int main(int argc, char** argv) {
  const int N = NFFT, R = 32, B = BLOCKS;
  const int U = argc > 1 ? atoi(argv[1]) : 32;      // units per launch
  const int reps = argc > 2 ? atoi(argv[2]) : 40;
  HostPlan hp; std::string err;
  #ifdef GNSSACQ_NO_PFA
  const std::function<bool(const HostPlan&, int)> use_pfa = [](const HostPlan&, int) { return false; };
#else
  const std::function<bool(const HostPlan&, int)> use_pfa = [](const HostPlan&, int) { return true; };
#endif
#ifdef SCHED1
  const std::vector<int> sched1 = {SCHED1, SCHED1B};
  const std::vector<int>* ps1 = &sched1;
#else
  const std::vector<int>* ps1 = nullptr;
#endif
  if (!make_plan(N, hp, err, 0, 0, ps1, nullptr, &use_pfa)) { printf("plan: %s\n", err.c_str()); return 1; }
  printf("N=%d plan %dx%d  U=%d B=%d\n", N, hp.N1, hp.N2, U, B);
  DevPlan dp{}; dp.N = N; dp.N1 = hp.N1; dp.N2 = hp.N2;
  fill_sub(hp.s1, dev(hp.tw1), dp.s1); fill_sub(hp.s2, dev(hp.tw2), dp.s2); dp.twm = dev(hp.twm);
  dp.twm_inv = hp.twm_inv.empty() ? dp.twm : dev(hp.twm_inv);
  dp.n1_of_pos = dev(hp.s1.n_of_pos); dp.n2_of_pos = dev(hp.s2.n_of_pos); dp.pos2_of_n = dev(hp.s2.pos_of_n);
  printf("prime-factor: N1 %d, N2 %d\n", (int)hp.s1.pfa, (int)hp.s2.pfa);
  if (!schedule_matches<SCOLS>(dp.s1) || !schedule_matches<SROWS>(dp.s2)) { printf("schedule mismatch\n"); return 1; }
  std::mt19937 rng(1);
  std::normal_distribution<float> nd(0.f, 1.f);
  const int Dc = (U + R - 1) / R;
  std::vector<float2> hX((size_t)Dc * B * N), hC((size_t)R * N);
  for (auto& v : hX) v = make_float2(nd(rng), nd(rng));
  for (auto& v : hC) v = make_float2(nd(rng), nd(rng));
  float2 *dX = dev(hX), *dC = dev(hC), *scr;
  CK(cudaMalloc(&scr, (size_t)U * B * N * sizeof(float2)));
  const int ntiles = (dp.N2 + kTileW - 1) / kTileW;
  Part *pa, *pb;
  const size_t np = (size_t)R * Dc * ntiles;
  CK(cudaMalloc(&pa, np * sizeof(Part))); CK(cudaMalloc(&pb, np * sizeof(Part)));
  CK(cudaMemset(pa, 0, np * sizeof(Part))); CK(cudaMemset(pb, 0, np * sizeof(Part)));
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));

  auto krows = k_corr_rows_s<SROWS>;
  auto kcols = k_corr_cols_s<SCOLS, kMulti>;
  auto kcolsp = k_corr_cols_p<SCOLS, kMulti>;
  auto krowsp = k_corr_rows_p<SROWS, RTHREADS>;
  const size_t smrp = rows_pipe_smem<SROWS>();
  CK(cudaFuncSetAttribute(krowsp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smrp));
  auto krowst = k_corr_rows_t<SROWS, TTHREADS, TMINCTAS>;
  auto kcolst = k_corr_cols_s<SCOLS, kMulti, CTHREADS, CMINCTAS, CSPLIT>;
  CK(cudaFuncSetAttribute(kcolst, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)hp.N1 * kTileW * (sizeof(float2) + (kMulti ? sizeof(float) : 0)))));
  const size_t smrt = rows_t_smem<SROWS>();
  CK(cudaFuncSetAttribute(krowst, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smrt));
  float2* scr2;
  CK(cudaMalloc(&scr2, (size_t)U * B * N * sizeof(float2)));
  CK(cudaMemset(scr2, 0xff, (size_t)U * B * N * sizeof(float2)));
  const size_t smr = rows_spec_smem<SROWS>(), smc = (size_t)dp.N1 * kTileW * (sizeof(float2) + (kMulti ? sizeof(float) : 0));
  const size_t smcp = cols_pipe_smem<SCOLS, kMulti>();
  CK(cudaFuncSetAttribute(krows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smr));
  CK(cudaFuncSetAttribute(kcols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smc));
  CK(cudaFuncSetAttribute(kcolsp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smcp));
  const float scale = 1.f / N;
  const int gridp = std::min(U * ntiles, 2 * nsm);
  auto run_rows = [&](cudaStream_t st) { krows<<<dim3((dp.N1 + kTileW - 1) / kTileW, B, U), kThreads, smr, st>>>(dp, dX, dC, R, B, 0, scr); };
  auto run_cols = [&](cudaStream_t st) { kcols<<<dim3(ntiles, U), kThreads, smc, st>>>(dp, scr, R, B, Dc, 0, 0, N, scale, ntiles, pa, nullptr); };
  auto run_colsp = [&](cudaStream_t st) { kcolsp<<<gridp, kThreads, smcp, st>>>(dp, scr, R, B, Dc, 0, 0, U, N, scale, ntiles, pb, nullptr); };

  const int nrt = (dp.N1 + kRowsTile8 - 1) / kRowsTile8;
  const int gridr = std::min(nrt * B * U, 2 * nsm);
  auto run_rowsp = [&](cudaStream_t st, float2* dst) { krowsp<<<gridr, RTHREADS, smrp, st>>>(dp, dX, dC, R, B, 0, U, dst); };
  auto run_rowst = [&](cudaStream_t st, float2* dst) { krowst<<<dim3(nrt, B, U), TTHREADS, smrt, st>>>(dp, dX, dC, R, B, 0, dst); };
  run_rows(0); run_cols(0); run_colsp(0);
  for (int pass = 0; pass < 2; ++pass) {
    CK(cudaMemset(scr2, 0xff, (size_t)U * B * N * sizeof(float2)));
    if (pass == 0) run_rowsp(0, scr2); else run_rowst(0, scr2);
    CK(cudaDeviceSynchronize());
    std::vector<float2> s1((size_t)U * B * N), s2((size_t)U * B * N);
    CK(cudaMemcpy(s1.data(), scr, s1.size() * sizeof(float2), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(s2.data(), scr2, s2.size() * sizeof(float2), cudaMemcpyDeviceToHost));
    size_t badr = 0;
    for (size_t i = 0; i < s1.size(); ++i) if (memcmp(&s1[i], &s2[i], sizeof(float2))) { if (badr < 5) printf("rows mismatch %zu: (%g,%g) vs (%g,%g)\n", i, s1[i].x, s1[i].y, s2[i].x, s2[i].y); ++badr; }
    printf("rows %s vs reference: %zu / %zu elements differ\n", pass ? "small-CTA" : "pipelined", badr, s1.size());
  }
  std::vector<Part> ha(np), hb(np);
  CK(cudaMemcpy(ha.data(), pa, np * sizeof(Part), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hb.data(), pb, np * sizeof(Part), cudaMemcpyDeviceToHost));
  size_t bad = 0;
  for (size_t i = 0; i < np; ++i) if (ha[i].key != hb[i].key || ha[i].sum != hb[i].sum) { if (bad < 5) printf("mismatch %zu: %llx %g vs %llx %g\n", i, ha[i].key, ha[i].sum, hb[i].key, hb[i].sum); ++bad; }
  printf("cols pipelined vs reference: %zu / %zu tiles differ\n", bad, np);

  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto time = [&](const char* name, auto&& f) {
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / reps;
    printf("%-28s %8.2f us/launch  %.3e cell-blocks/s\n", name, us, (double)U * B * N / us * 1e6);
  };
  time("rows (spec)", [&] { run_rows(0); });
  time("cols (spec)", [&] { run_cols(0); });
  time("cols (pipelined)", [&] { run_colsp(0); });
  time("rows (small CTAs)", [&] { run_rowst(0, scr); });
  time("cols (small CTAs)", [&] { kcolst<<<dim3(ntiles, U), CTHREADS, smc, 0>>>(dp, scr, R, B, Dc, 0, 0, N, scale, ntiles, pa, nullptr); });
  time("rows+cols (small CTAs)", [&] { run_rowst(0, scr); kcolst<<<dim3(ntiles, U), CTHREADS, smc, 0>>>(dp, scr, R, B, Dc, 0, 0, N, scale, ntiles, pa, nullptr); });
  time("rows (pipelined)", [&] { run_rowsp(0, scr); });
  time("rows+cols (both pipelined)", [&] { run_rowsp(0, scr); run_colsp(0); });
  time("rows+cols (spec)", [&] { run_rows(0); run_cols(0); });
  time("rows+cols (pipelined cols)", [&] { run_rows(0); run_colsp(0); });
  CK(cudaGetLastError());
  return bad != 0;
}
