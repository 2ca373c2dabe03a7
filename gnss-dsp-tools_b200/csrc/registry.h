// Lookup of the plan-specialised kernels by radix schedule. The template instantiations behind
// these functions are compiled in separate translation units (registry.cu, one part per nvcc
// invocation, see __graft_entry__.build) so that the library builds in parallel.
#pragma once
#include "fft_core.cuh"
#include "kernels.cuh"
#include "async_copy.cuh"

namespace acq {

typedef void (*corr_rows_fn)(DevPlan, const float2*, const float2*, int, int, int, float2*);
typedef void (*corr_cols_fn)(DevPlan, const float2*, int, int, int, int, int, int, float, int, Part*, float*, unsigned*);
typedef void (*fwd_cols_fn)(DevPlan, const float2*, const float*, const double*, const float2*, int, int, float2*);
typedef void (*fwd_rows_fn)(DevPlan, float2*);

// 256-thread kernels (kernels_spec.cuh); nullptr when the schedule has no specialisation
fwd_cols_fn find_fwd_cols_kernel(const SubPlan& s1, int src);
fwd_rows_fn find_fwd_rows_kernel(const SubPlan& s2);
corr_rows_fn find_rows_kernel(const SubPlan& s2, bool gt = false);   // gt: coprime split, no four-step twiddle
corr_cols_fn find_cols_kernel(const SubPlan& s1, bool multi);

// small-CTA kernels (kernels_small.cuh): rows grid = (ceil(N1/8), B, units) with `smem` bytes,
// columns grid and shared memory as the 256-thread kernel
struct RowsSmall { corr_rows_fn fn; int threads; size_t smem; };
struct ColsSmall { corr_cols_fn fn; int threads; };
constexpr int kRowsSmallTile = 8;
RowsSmall find_rows_small(const SubPlan& s2, bool gt = false);
ColsSmall find_cols_small(const SubPlan& s1, bool multi);

// Copy-engine-fed pair for coprime plans with two-stage schedules (kernels_v3.cuh). `variant`
// picks among the instantiated tile shapes (0 = default, measured fastest; A/B through the
// "v3_rows" / "v3_cols" options); fn == nullptr when the schedule has none.
typedef void (*rows_v3_fn)(DevPlan, const float2*, const float2*, ChunkV3, int, float2*);
typedef void (*cols_v3_fn)(DevPlan, const TensorMap, int, const int*, ChunkV3, int, int, int, int, float, int, Part*, float*, unsigned*);
struct RowsV3 { rows_v3_fn fn; int threads, T, ctas_per_sm; size_t smem; int RA, RB, PB; };
struct ColsV3 { cols_v3_fn fn; int threads, CW, ctas_per_sm; size_t smem; int parts_per_tile; };   // parts per (unit, tile): 1, or one per warp
RowsV3 find_rows_v3(const SubPlan& s2, int variant);
ColsV3 find_cols_v3(const SubPlan& s1, bool multi, bool dump, int variant);   // dump: also writes the q grid (tests)
// forward rows kernel in the two-role structure (k_fwd_rows_v6): grid (row tiles of T rows, splits of the transform list)
typedef void (*fwd_rows_v6_fn)(DevPlan, float2*, int);
struct FwdRowsV6 { fwd_rows_v6_fn fn; int threads, T, ctas_per_sm; size_t smem; };
FwdRowsV6 find_fwd_rows_v6(const SubPlan& s2);

// Fused persistent correlate kernel (kernels_fused.cuh): both task types in one launch per Doppler chunk.
typedef void (*fused_fn)(DevPlan, const TensorMap, const int*, FusedJob, FusedSync, const float2*, const float2*, float2*, Part*, float*, unsigned*);
struct FusedKernel { fused_fn fn; int threads, T, CW, ctas_per_sm; size_t smem; int RA, RB, PB; };
FusedKernel find_fused(const SubPlan& s1, const SubPlan& s2, bool multi, bool dump);

}  // namespace acq
