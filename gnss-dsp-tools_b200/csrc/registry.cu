// Instantiations of the plan-specialised kernels, split into parts (-DGNSSACQ_REG_PART=k) so
// that nvcc compiles them in parallel. Schedules are exactly those fft_plan.h::make_subplan
// emits for the FFT lengths GNSS receivers use; each is checked against the runtime plan.
#include "registry.h"
#include "kernels_small.cuh"
#include "kernels_v3.cuh"
#include "kernels_fused.cuh"

namespace acq {

#define COLS_LIST_0(T) T(S128) T(S256) T(S320) T(S165)
#define COLS_LIST_1(T) T(S220) T(S372b) T(S200) T(S248)
#define COLS_LIST_2(T) T(S496) T(S186) T(S279) T(S341)

corr_cols_fn find_cols_part0(const SubPlan&, bool);
corr_cols_fn find_cols_part1(const SubPlan&, bool);
corr_cols_fn find_cols_part2(const SubPlan&, bool);
corr_cols_fn find_cols_small_part0(const SubPlan&, bool);
corr_cols_fn find_cols_small_part1(const SubPlan&, bool);
corr_cols_fn find_cols_small_part2(const SubPlan&, bool);
constexpr int kColsSmallThreads = 128;

#if GNSSACQ_REG_PART == 0
fwd_cols_fn find_fwd_cols_kernel(const SubPlan& s1, int src) {
#define TRY(S) if (schedule_matches<S>(s1)) return src == 0 ? k_fwd_cols_s<S, 0> : k_fwd_cols_s<S, 1>;
  TRY(S128) TRY(S256) TRY(S320) TRY(S165) TRY(S220) TRY(S372b) TRY(S200) TRY(S248) TRY(S496) TRY(S186) TRY(S279) TRY(S341)
#undef TRY
  return nullptr;
}
fwd_rows_fn find_fwd_rows_kernel(const SubPlan& s2) {
#define TRY(S) if (schedule_matches<S>(s2)) return k_fwd_rows_s<S>;
  TRY(S128) TRY(S256) TRY(S512) TRY(S320) TRY(S186) TRY(S279) TRY(S440) TRY(S250) TRY(S165) TRY(S220) TRY(S480) TRY(S90)
#undef TRY
  return nullptr;
}
corr_cols_fn find_cols_kernel(const SubPlan& s1, bool multi) {
  if (corr_cols_fn f = find_cols_part0(s1, multi)) return f;
  if (corr_cols_fn f = find_cols_part1(s1, multi)) return f;
  return find_cols_part2(s1, multi);
}
ColsSmall find_cols_small(const SubPlan& s1, bool multi) {
  corr_cols_fn f = find_cols_small_part0(s1, multi);
  if (!f) f = find_cols_small_part1(s1, multi);
  if (!f) f = find_cols_small_part2(s1, multi);
  return ColsSmall{f, f ? kColsSmallThreads : 0};
}
#elif GNSSACQ_REG_PART == 1
corr_rows_fn find_rows_kernel(const SubPlan& s2, bool gt) {
  if (gt) {                              // coprime splits (no four-step twiddle): the lengths that have one
#define TRY(S) if (schedule_matches<S>(s2)) return k_corr_rows_s<S, true>;
    TRY(S220) TRY(S480) TRY(S90)
#undef TRY
    return nullptr;
  }
#define TRY(S) if (schedule_matches<S>(s2)) return k_corr_rows_s<S>;
  TRY(S128) TRY(S256) TRY(S512) TRY(S320) TRY(S186) TRY(S279) TRY(S440) TRY(S250) TRY(S165) TRY(S220)
#undef TRY
  return nullptr;
}
// threads / CTAs per SM: enough registers for the widest in-register butterfly of the schedule.
// Listed = measured faster than the 256-thread kernel on B200 (tools/bench_configs.py small_ctas=0..3,
// profiles/README.md r02): 2-20 %; 256 = 16*16 lost 2-7 % and stays on the 256-thread kernel.
RowsSmall find_rows_small(const SubPlan& s2, bool gt) {
  static_assert(kRowsSmallTile == kRowsTile8, "row tile");
  if (gt) {
#define TRY(S, T, C) if (schedule_matches<S>(s2)) return RowsSmall{k_corr_rows_t<S, T, C, true>, T, rows_t_smem<S>()};
    TRY(S220, 128, 4) TRY(S480, 128, 4) TRY(S90, 128, 6)
#undef TRY
    return RowsSmall{nullptr, 0, 0};
  }
#define TRY(S, T, C) if (schedule_matches<S>(s2)) return RowsSmall{k_corr_rows_t<S, T, C>, T, rows_t_smem<S>()};
  TRY(S440, 192, 5) TRY(S128, 128, 6) TRY(S512, 128, 6) TRY(S320, 128, 6) TRY(S220, 128, 4) TRY(S250, 128, 4)
#undef TRY
  return RowsSmall{nullptr, 0, 0};
}
#elif GNSSACQ_REG_PART >= 2 && GNSSACQ_REG_PART <= 4
#define TRY(S) if (schedule_matches<S>(s1)) return multi ? k_corr_cols_s<S, true> : k_corr_cols_s<S, false>;
#if GNSSACQ_REG_PART == 2
corr_cols_fn find_cols_part0(const SubPlan& s1, bool multi) { COLS_LIST_0(TRY) return nullptr; }
#elif GNSSACQ_REG_PART == 3
corr_cols_fn find_cols_part1(const SubPlan& s1, bool multi) { COLS_LIST_1(TRY) return nullptr; }
#else
corr_cols_fn find_cols_part2(const SubPlan& s1, bool multi) { COLS_LIST_2(TRY) return nullptr; }
#endif
#undef TRY
#elif GNSSACQ_REG_PART >= 5 && GNSSACQ_REG_PART <= 7
// Listed = measured faster than the 256-thread kernel (same measurements): every schedule with a
// radix-31 stage (3-10 %), 128 = 8*16 (17 %), 200 = 10*20 (2 %), 256 = 16*16 for one block only
// (+9 %; -5 % with the non-coherent accumulator in shared memory). 320 = 5*8*8 lost 8 %.
#define TRY(S) if (schedule_matches<S>(s1)) return multi ? k_corr_cols_s<S, true, kColsSmallThreads, 4, false> : k_corr_cols_s<S, false, kColsSmallThreads, 4, false>;
#if GNSSACQ_REG_PART == 5
corr_cols_fn find_cols_small_part0(const SubPlan& s1, bool multi) {
  TRY(S128)
  if (!multi && schedule_matches<S256>(s1)) return k_corr_cols_s<S256, false, kColsSmallThreads, 4, false>;
  return nullptr;
}
#elif GNSSACQ_REG_PART == 6
corr_cols_fn find_cols_small_part1(const SubPlan& s1, bool multi) { TRY(S372b) TRY(S200) TRY(S248) return nullptr; }
#else
corr_cols_fn find_cols_small_part2(const SubPlan& s1, bool multi) { COLS_LIST_2(TRY) return nullptr; }
#endif
#undef TRY
#elif GNSSACQ_REG_PART == 8
// rows: (tile rows, threads, CTAs per SM, ring slots)
RowsV3 find_rows_v3(const SubPlan& s2, int variant) {
#define TRY(S, V, T, TH, C, XB)                                                                                   \
  if (variant == V && schedule_matches<S>(s2))                                                                    \
    return RowsV3{k_corr_rows_v3<S, T, TH, C, XB>, TH, T, C, rows_v3_smem<S, T, XB>(), S::radix(0), S::radix(1), v3_pitch(S::radix(1))};
  TRY(S480, 5, 4, 128, 4, 1) TRY(S480, 1, 8, 256, 1, 2) TRY(S480, 2, 8, 256, 2, 1) TRY(S480, 3, 4, 128, 3, 2)
  TRY(S220, 0, 8, 160, 3, 2) TRY(S220, 1, 8, 160, 4, 1)
  TRY(S90, 0, 8, 96, 8, 1) TRY(S90, 1, 8, 96, 6, 2)
#undef TRY
  // balanced kernel (the two halves of the CTA alternate on stage B): 4 warps, stage B = 2 warps
  if (variant == 4 && schedule_matches<S480>(s2))
    return RowsV3{k_corr_rows_v4<S480, 4, 128, 3>, 128, 4, 3, rows_v4_smem<S480, 4>(), S480::radix(0), S480::radix(1), v3_pitch(S480::radix(1))};
  // (forward counterpart below)
  // default for 480: two roles, warp 0 = stage A of a 2-row tile, warp 1 = stage B + bulk store, no block barrier
  // (r04d: correlate stage 1.63 -> 1.59 ms per config-2 step against the 4-row x 128-thread kernel, now variant 5)
  if (variant == 0 && schedule_matches<S480>(s2))
    return RowsV3{k_corr_rows_v6<S480, 7>, 64, 2, 7, rows_v6_smem<S480>(), S480::radix(0), S480::radix(1), v3_pitch(S480::radix(1))};
  return RowsV3{nullptr, 0, 0, 0, 0, 0, 0, 0};
}
FwdRowsV6 find_fwd_rows_v6(const SubPlan& s2) {
  if (schedule_matches<S480>(s2)) return FwdRowsV6{k_fwd_rows_v6<S480, 7>, 64, 2, 7, fwd_rows_v6_smem<S480>()};
  return FwdRowsV6{nullptr, 0, 0, 0, 0};
}
#elif GNSSACQ_REG_PART == 9
// cols: (tile columns, threads, CTAs per SM)
ColsV3 find_cols_v3(const SubPlan& s1, bool multi, bool dump, int variant) {
#define PICK(S, CW, TH, C, M, DUMP) ColsV3{k_corr_cols_v3<S, M, DUMP, CW, TH, C>, TH, CW, C, cols_v3_smem<S, M, CW>(), 1}
#define PICK0(S, CW, TH, C, M, DUMP) ColsV3{k_corr_cols_v3<S, M, DUMP, CW, TH, C, false>, TH, CW, C, cols_v3_smem<S, M, CW>(), 1}
// multi-block searches with the non-coherent sums in registers (no q array in shared memory); single-block calls get the plain kernel
#define PICKQ(S, CW, TH, C, M, DUMP) ColsV3{k_corr_cols_v3<S, M, DUMP, CW, TH, C, true, M>, TH, CW, C, cols_v3_smem<S, false, CW>(), 1}
#define PICK5(S, CW, TH, C, M, DUMP) ColsV3{k_corr_cols_v5<S, M, DUMP, CW, TH, C>, TH, CW, C, cols_v5_smem<S, M, CW>(), 1}
#define TRY(P, S, V, CW, TH, C)                                                                                   \
  if (variant == V && schedule_matches<S>(s1))                                                                    \
    return multi ? (dump ? P(S, CW, TH, C, true, true) : P(S, CW, TH, C, true, false))                            \
                 : (dump ? P(S, CW, TH, C, false, true) : P(S, CW, TH, C, false, false));
  // default: 8-column tiles x 96 threads x 5 CTAs per SM; multi-block searches keep their non-coherent sums in registers
  // (4 CTAs per SM, no q array in shared memory: +1-2 % on 61380 x 20 blocks and 30690 x 20, r04)
  if (variant == 0 && multi && schedule_matches<S341>(s1)) return dump ? PICKQ(S341, 8, 96, 4, true, true) : PICKQ(S341, 8, 96, 4, true, false);
  if (variant == 0 && multi && schedule_matches<S279>(s1)) return dump ? PICKQ(S279, 8, 96, 4, true, true) : PICKQ(S279, 8, 96, 4, true, false);
  TRY(PICK, S341, 0, 8, 96, 5) TRY(PICK, S341, 1, 16, 192, 2) TRY(PICK, S341, 2, 16, 256, 2)
  TRY(PICK, S341, 6, 8, 96, 5) TRY(PICK, S279, 6, 8, 96, 5)            // A/B: non-coherent sums in shared memory
  TRY(PICKQ, S279, 7, 8, 96, 5)
  TRY(PICK0, S341, 5, 8, 96, 5)                                        // A/B: per-tile chores on thread 0 instead of the last warp
  TRY(PICK5, S341, 3, 8, 96, 6) TRY(PICK5, S341, 4, 8, 96, 7)          // one tile slot, next copy issued behind the radix-31 loads
  TRY(PICK, S279, 0, 8, 96, 5) TRY(PICK, S279, 1, 16, 160, 3)
  TRY(PICK5, S279, 3, 8, 96, 6)
#undef TRY
#undef PICK
#undef PICK5
#undef PICK0
#undef PICKQ
  return ColsV3{nullptr, 0, 0, 0, 0, 0};
}
#elif GNSSACQ_REG_PART == 10
FusedKernel find_fused(const SubPlan& s1, const SubPlan& s2, bool multi, bool dump) {
#define PICK(SC, SR, T, CW, TH, C, M, DUMP) \
  FusedKernel{k_corr_fused<SR, SC, T, CW, M, DUMP, TH, C>, TH, T, CW, C, fused_smem<SR, SC, T, CW, M>(), SR::radix(0), SR::radix(1), v3_pitch(SR::radix(1))}
#define TRY(SC, SR, T, CW, TH, C)                                                                                  \
  if (schedule_matches<SC>(s1) && schedule_matches<SR>(s2))                                                        \
    return multi ? (dump ? PICK(SC, SR, T, CW, TH, C, true, true) : PICK(SC, SR, T, CW, TH, C, true, false))       \
                 : (dump ? PICK(SC, SR, T, CW, TH, C, false, true) : PICK(SC, SR, T, CW, TH, C, false, false));
  TRY(S341, S480, 4, 8, 128, 4) TRY(S279, S220, 8, 8, 160, 3)
#undef TRY
#undef PICK
  return FusedKernel{nullptr, 0, 0, 0, 0, 0, 0, 0, 0};
}
#else
#error "GNSSACQ_REG_PART must be 0..10"
#endif

}  // namespace acq
