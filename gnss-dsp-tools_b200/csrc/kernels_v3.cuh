// Copy-engine-fed correlate kernels for coprime (Good-Thomas) plans whose two tile transforms are
// two-stage prime-factor schedules: 163680 = 341 x 480 (31*11, 15*32), 61380 = 279 x 220 (31*9, 11*20).
//
// Why: ncu on the register-loading kernels (kernels_small.cuh) shows them latency-bound — issue
// slots 34-40 % busy, 53 % of the rows kernel's stall samples waiting on global loads — not
// bandwidth-bound (profiles/README.md, r03a). Here the tiles are moved by the copy engine
// (cp.async.bulk / cp.async.bulk.tensor + mbarrier transaction counts) while the butterflies of the
// previous tile run, so no warp ever waits on a global load with data in registers:
//
//   rows kernel  one CTA = (tile of T rows, replica r); it keeps its slice of the replica spectrum
//                C[r] in registers and walks the (Doppler, block) pairs of the chunk: the spectra
//                tiles X[d][b] stream through a shared-memory ring (1-D bulk copies, they are
//                contiguous), stage A = multiply + radix-RA butterflies over the stride-RB digit
//                (lanes along the contiguous digit), exchange through a padded tile, stage B =
//                radix-RB butterflies over 16-byte accesses in place, and the finished tile leaves
//                as ONE bulk store. No four-step twiddle (coprime split), no stage twiddles
//                (coprime radices), one shared-memory exchange.
//   cols kernel  persistent CTAs; tile = N1 x CW scratch columns fetched with one 3-D tensor-map
//                copy into a two-slot ring; first stage in place, then the radix-31 stage fused with
//                |.|, the non-coherent sum and the peak search (cols_last_stage).
//
// Scratch layout (private to this pair): row p1 of slot s holds RA groups of PB = RB + pad
// elements, element (a', b') at ((s*N1 + p1)*RA + a')*PB + b'; pad elements are never read as data
// (col_lag_p marks them -1). The pad makes stage B's 16-byte accesses conflict-free and lets the
// bulk store copy the tile verbatim.
#pragma once
#include "async_copy.cuh"
#include "kernels_small.cuh"

namespace acq {

// pitch (elements) of one RB-group in the exchange tile: a multiple of 2 (16-byte accesses) with
// pitch/2 odd, so 8 consecutive groups start in 8 different 16-byte bank groups
__host__ __device__ constexpr int v3_pitch(int RB) {
  int p = RB + (RB & 1);
  while ((p / 2) % 2 == 0) p += 2;
  return p;
}

// =========================================================================== rows kernel
template <class S, int T, int XBUFS> __host__ __device__ constexpr size_t rows_v3_smem() {
  return (size_t)XBUFS * T * S::F * sizeof(float2) + 2 * (size_t)T * S::radix(0) * v3_pitch(S::radix(1)) * sizeof(float2) + XBUFS * 8;
}

// Stage A of one rows tile: this thread's butterfly (row ra, contiguous digit b) — multiply the
// spectra tile by the replica values held in registers, radix-RA inverse butterfly over the
// stride-RB digit, scatter into the padded exchange tile.
template <class S>
__device__ __forceinline__ void rows_v3_stage_a(const float2* __restrict__ xtile, float2* __restrict__ et, const float2* c, int ra, int b) {
  constexpr int N2 = S::F, RA = S::radix(0), RB = S::radix(1), PB = v3_pitch(RB), NP = RA * PB;
  const float2* xp = xtile + ra * N2 + b;
  float2 y[RA];
#pragma unroll
  for (int a = 0; a < RA; ++a) y[a] = cmulc(c[a], xp[a * RB]);
  inv_dft<RA>(y);
  float2* ep = et + ra * NP + b;
#pragma unroll
  for (int a = 0; a < RA; ++a) ep[a * PB] = y[a];
}
// Stage B: radix-RB inverse butterfly of group `grp` (= row * RA + a') in place, 16-byte accesses.
template <class S>
__device__ __forceinline__ void rows_v3_stage_b(float2* __restrict__ et, int grp) {
  constexpr int RB = S::radix(1), PB = v3_pitch(RB);
  float4* p4 = reinterpret_cast<float4*>(et + grp * PB);
  float2 v[RB];
#pragma unroll
  for (int q = 0; q < RB / 2; ++q) { const float4 t = p4[q]; v[2 * q] = make_float2(t.x, t.y); v[2 * q + 1] = make_float2(t.z, t.w); }
  inv_dft<RB>(v);
#pragma unroll
  for (int q = 0; q < RB / 2; ++q) p4[q] = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
}

// grid = (row tiles, Rc, splits of the pair list); THREADS >= T * RB. X: [Dc][B][N], C: [R][N] (position order), scratch as above.
template <class S, int T, int THREADS, int MINCTAS, int XBUFS>
__global__ void __launch_bounds__(THREADS, MINCTAS)
k_corr_rows_v3(DevPlan pl, const float2* __restrict__ X, const float2* __restrict__ C, ChunkV3 ck, int B,
               float2* __restrict__ scratch) {
  GNSSACQ_DYN_SMEM(float2, smem);
  static_assert(S::NS == 2 && S::kPfa, "two coprime stages");
  constexpr int N2 = S::F, RA = S::radix(0), RB = S::radix(1), PB = v3_pitch(RB), NP = RA * PB;
  static_assert(RB % 2 == 0 && THREADS >= T * RB && THREADS >= T * RA, "thread mapping");
  constexpr int XT = T * N2, ET = T * NP;                     // float2 per X buffer / exchange buffer
  float2* xbuf = smem;
  float2* ebuf = smem + XBUFS * XT;
  unsigned long long* full = reinterpret_cast<unsigned long long*>(ebuf + 2 * ET);
  const int N = pl.N, N1 = pl.N1;
  const int row0 = blockIdx.x * T;
  const int nrows = imin(T, N1 - row0);
  const int r = ck.r0 + blockIdx.y;
  // (Doppler, block) pairs, block fastest; grid.z splits the list so that short chunks still fill the GPU
  const int nall = ck.G * B, per = (nall + gridDim.z - 1) / gridDim.z;
  const int it0 = blockIdx.z * per, nit = imin(per, nall - it0);
  const int tid = threadIdx.x;
  const unsigned xbytes = (unsigned)nrows * N2 * sizeof(float2), ebytes = (unsigned)nrows * NP * sizeof(float2);

  auto issue = [&](int it) {                                   // thread 0: spectra tile of pair `it` -> ring slot it % XBUFS
    const int xs = it % XBUFS;
    const int g = it0 + it;
    const float2* src = X + ((long long)(ck.dd0 + g / B) * B + g % B) * N + (long long)row0 * N2;
    mbar_arrive_expect(&full[xs], xbytes);
    bulk_g2s(xbuf + xs * XT, src, xbytes, &full[xs]);
  };
  if (tid == 0) {
    for (int s = 0; s < XBUFS; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
    for (int it = 0; it < XBUFS && it < nit; ++it) issue(it);
  }
  // stage-A butterfly of this thread: row ra, contiguous digit b; its replica-spectrum values stay in registers
  const int ra = tid / RB, b = tid - ra * RB;
  const bool act_a = tid < T * RB && ra < nrows;
  float2 c[RA];
  if (act_a) {
    const float2* cp = C + (long long)r * N + (long long)(row0 + ra) * N2 + b;
#pragma unroll
    for (int a = 0; a < RA; ++a) c[a] = __ldg(&cp[a * RB]);
  }
  const bool act_b = tid < nrows * RA;                         // stage-B butterfly: group tid = row * RA + a'
  __syncthreads();                                             // barrier initialisation visible to every thread

  for (int it = 0; it < nit; ++it) {
    const int xs = it % XBUFS, e = it & 1;
    float2* et = ebuf + e * ET;
    mbar_wait(&full[xs], (unsigned)(it / XBUFS) & 1u);
    if (act_a) rows_v3_stage_a<S>(xbuf + xs * XT, et, c, ra, b);
    __syncthreads();                                           // exchange tile complete; ring slot xs drained
    if (tid == 0 && it + XBUFS < nit) { fence_async_smem(); issue(it + XBUFS); }
    if (act_b) {
      rows_v3_stage_b<S>(et, tid);
      fence_async_smem();                                      // these writes are read by the bulk store below
    }
    if (tid == 0) bulk_wait_read<0>();                         // the store of pair it-1 has drained the other exchange buffer
    __syncthreads();
    if (tid == 0) {
      const int g = it0 + it;
      const int slot = v3_slot(ck, B, r, ck.dd0 + g / B, g % B);
      bulk_s2g(scratch + ((long long)slot * N1 + row0) * NP, et, ebytes);
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait_all<0>();                            // shared memory must outlive the last store
}

// =========================================================================== rows kernel, balanced
// Stage B needs T*RA threads — half of the CTA for 480 = 15 * 32 — so in k_corr_rows_v3 half of the warps
// sit at the block barrier while the others run the radix-32 butterflies (ncu r03d: 35 % of the stall
// samples are barrier waits). Here the two halves of the CTA take turns: pair (it mod 2) runs stage B of
// (Doppler, block) pair `it` while the other half already multiplies and transforms pair it+1 into the
// second exchange buffer. Block barriers become mbarriers:
//   xfull[s]  spectra tile of ring slot s landed (copy-engine transaction count)
//   efull[e]  every thread has written its stage-A outputs into exchange buffer e (count THREADS)
//   efree[e]  the bulk store that read exchange buffer e has drained it (arrived by the half that issued it)
// Over two pairs every warp does A, A, B. Same arithmetic as k_corr_rows_v3: bit-identical scratch.
template <class S, int T> __host__ __device__ constexpr size_t rows_v4_smem() {
  return 2 * (size_t)T * S::F * sizeof(float2) + 2 * (size_t)T * S::radix(0) * v3_pitch(S::radix(1)) * sizeof(float2) + 6 * 8;
}
template <class S, int T, int THREADS, int MINCTAS>
__global__ void __launch_bounds__(THREADS, MINCTAS)
k_corr_rows_v4(DevPlan pl, const float2* __restrict__ X, const float2* __restrict__ C, ChunkV3 ck, int B,
               float2* __restrict__ scratch) {
  GNSSACQ_DYN_SMEM(float2, smem);
  static_assert(S::NS == 2 && S::kPfa, "two coprime stages");
  constexpr int N2 = S::F, RA = S::radix(0), RB = S::radix(1), PB = v3_pitch(RB), NP = RA * PB;
  constexpr int GT = (T * RA + 31) / 32 * 32;                 // threads of one stage-B half
  static_assert(RB % 2 == 0 && THREADS >= T * RB && THREADS == 2 * GT, "two halves, each large enough for stage B");
  constexpr int XT = T * N2, ET = T * NP;
  float2* xbuf = smem;
  float2* ebuf = smem + 2 * XT;
  unsigned long long* xfull = reinterpret_cast<unsigned long long*>(ebuf + 2 * ET);
  unsigned long long* efull = xfull + 2;
  unsigned long long* efree = xfull + 4;
  const int N = pl.N, N1 = pl.N1;
  const int row0 = blockIdx.x * T;
  const int nrows = imin(T, N1 - row0);
  const int r = ck.r0 + blockIdx.y;
  const int nall = ck.G * B, per = (nall + gridDim.z - 1) / gridDim.z;
  const int it0 = blockIdx.z * per, nit = imin(per, nall - it0);
  const int tid = threadIdx.x, half = tid / GT, htid = tid - half * GT;
  const bool leader = htid == 0;
  const unsigned xbytes = (unsigned)nrows * N2 * sizeof(float2), ebytes = (unsigned)nrows * NP * sizeof(float2);

  auto issue = [&](int it) {                                   // spectra tile of pair `it` -> ring slot it & 1
    const int g = it0 + it;
    const float2* src = X + ((long long)(ck.dd0 + g / B) * B + g % B) * N + (long long)row0 * N2;
    mbar_arrive_expect(&xfull[it & 1], xbytes);
    bulk_g2s(xbuf + (it & 1) * XT, src, xbytes, &xfull[it & 1]);
  };
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&xfull[s], 1); mbar_init(&efull[s], THREADS); mbar_init(&efree[s], 1); }
    mbar_fence_init();
    for (int it = 0; it < 2 && it < nit; ++it) issue(it);
  }
  const int ra = tid / RB, b = tid - ra * RB;
  const bool act_a = tid < T * RB && ra < nrows;
  float2 c[RA];
  if (act_a) {
    const float2* cp = C + (long long)r * N + (long long)(row0 + ra) * N2 + b;
#pragma unroll
    for (int a = 0; a < RA; ++a) c[a] = __ldg(&cp[a * RB]);
  }
  __syncthreads();                                             // barrier initialisation visible to every thread

  for (int it = 0; it < nit; ++it) {
    const int e = it & 1;
    const unsigned ph = (unsigned)(it >> 1) & 1u;
    float2* et = ebuf + e * ET;
    mbar_wait(&xfull[e], ph);
    if (it >= 2) mbar_wait(&efree[e], ph ^ 1u);                // the store of pair it-2 has drained this buffer
    if (act_a) rows_v3_stage_a<S>(xbuf + e * XT, et, c, ra, b);
    mbar_arrive(&efull[e]);
    if (half == e) {
      mbar_wait(&efull[e], ph);                                // exchange buffer complete, ring slot e drained
      if (leader && it + 2 < nit) { fence_async_smem(); issue(it + 2); }
      if (htid < nrows * RA) { rows_v3_stage_b<S>(et, htid); fence_async_smem(); }
      named_bar_sync(1 + half, GT);
      if (leader) {
        const int g = it0 + it;
        const int slot = v3_slot(ck, B, r, ck.dd0 + g / B, g % B);
        bulk_s2g(scratch + ((long long)slot * N1 + row0) * NP, et, ebytes);
        bulk_commit();
      }
    } else if (leader && it >= 1) {
      bulk_wait_read<0>();                                     // my store of pair it-1 (issued one stage A ago) has read its buffer
      mbar_arrive(&efree[half]);
    }
  }
  if (leader) bulk_wait_all<0>();                              // shared memory must outlive the last stores
}

// =========================================================================== rows kernel, two roles
// In k_corr_rows_v3 stage B (T*RA butterflies of ~400 instructions) runs on half of the threads stage A
// (T*RB butterflies of ~190) uses, between two block barriers: a third of the warp-time of a CTA is spent
// waiting (ncu r04a: 34 % barrier stalls), and k_corr_rows_v4's alternating halves do not change that sum
// (each half still waits for the other half's share of stage A). Here the two stages belong to two warps
// for good: warp 0 multiplies and runs stage A of BOTH rows of a 2-row tile (two butterflies per lane),
// warp 1 runs stage B (2 * 15 = 30 lanes) and issues the bulk store — equal instruction counts per
// (Doppler, block) pair, and A(it+1) overlaps B(it) through the two exchange buffers. No block barrier:
//   xfull[s]  spectra tile of ring slot s landed (copy-engine transaction count)
//   efull[e]  the 32 lanes of warp 0 have written their stage-A outputs into exchange buffer e
//   efree[e]  the bulk store that read exchange buffer e has drained it (lane 0 of warp 1, at the start of the next pair)
// Same arithmetic as k_corr_rows_v3: bit-identical scratch.
template <class S> __host__ __device__ constexpr size_t rows_v6_smem() {
  return 2 * (size_t)2 * S::F * sizeof(float2) + 2 * (size_t)2 * S::radix(0) * v3_pitch(S::radix(1)) * sizeof(float2) + 6 * 8;
}
template <class S, int MINCTAS>
__global__ void __launch_bounds__(64, MINCTAS)
k_corr_rows_v6(DevPlan pl, const float2* __restrict__ X, const float2* __restrict__ C, ChunkV3 ck, int B,
               float2* __restrict__ scratch) {
  GNSSACQ_DYN_SMEM(float2, smem);
  static_assert(S::NS == 2 && S::kPfa, "two coprime stages");
  constexpr int T = 2;
  constexpr int N2 = S::F, RA = S::radix(0), RB = S::radix(1), PB = v3_pitch(RB), NP = RA * PB;
  static_assert(RB == 32 && T * RA <= 32, "one lane per stage-A column, stage B within one warp");
  constexpr int XT = T * N2, ET = T * NP;
  float2* xbuf = smem;
  float2* ebuf = smem + 2 * XT;
  unsigned long long* xfull = reinterpret_cast<unsigned long long*>(ebuf + 2 * ET);
  unsigned long long* efull = xfull + 2;
  unsigned long long* efree = xfull + 4;
  const int N = pl.N, N1 = pl.N1;
  const int row0 = blockIdx.x * T;
  const int nrows = imin(T, N1 - row0);
  const int r = ck.r0 + blockIdx.y;
  const int nall = ck.G * B, per = (nall + gridDim.z - 1) / gridDim.z;
  const int it0 = blockIdx.z * per, nit = imin(per, nall - it0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned xbytes = (unsigned)nrows * N2 * sizeof(float2), ebytes = (unsigned)nrows * NP * sizeof(float2);

  auto issue = [&](int it) {                                   // spectra tile of pair `it` -> ring slot it & 1
    const int g = it0 + it;
    const float2* src = X + ((long long)(ck.dd0 + g / B) * B + g % B) * N + (long long)row0 * N2;
    mbar_arrive_expect(&xfull[it & 1], xbytes);
    bulk_g2s(xbuf + (it & 1) * XT, src, xbytes, &xfull[it & 1]);
  };
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&xfull[s], 1); mbar_init(&efull[s], 32); mbar_init(&efree[s], 1); }
    mbar_fence_init();
    for (int it = 0; it < 2 && it < nit; ++it) issue(it);
  }
  __syncthreads();                                             // barrier initialisation visible to both warps

  if (warp == 0) {
    // stage A: lane = contiguous digit b, both rows; the replica-spectrum values stay in registers
    float2 c0[RA], c1[RA];
    const float2* cp = C + (long long)r * N + (long long)row0 * N2 + lane;
#pragma unroll
    for (int a = 0; a < RA; ++a) c0[a] = __ldg(&cp[a * RB]);
#pragma unroll
    for (int a = 0; a < RA; ++a) c1[a] = nrows > 1 ? __ldg(&cp[N2 + a * RB]) : make_float2(0.f, 0.f);
    for (int it = 0; it < nit; ++it) {
      const int e = it & 1;
      const unsigned ph = (unsigned)(it >> 1) & 1u;
      float2* et = ebuf + e * ET;
      mbar_wait(&xfull[e], ph);
      if (it >= 2) mbar_wait(&efree[e], ph ^ 1u);              // the store of pair it-2 has drained this buffer
      rows_v3_stage_a<S>(xbuf + e * XT, et, c0, 0, lane);
      if (nrows > 1) rows_v3_stage_a<S>(xbuf + e * XT, et, c1, 1, lane);
      __syncwarp();                                            // every lane is done with ring slot e
      if (lane == 0 && it + 2 < nit) { fence_async_smem(); issue(it + 2); }
      mbar_arrive(&efull[e]);
    }
  } else {
    const bool act_b = lane < nrows * RA;                      // stage-B butterfly: group lane = row * RA + a'
    for (int it = 0; it < nit; ++it) {
      const int e = it & 1;
      const unsigned ph = (unsigned)(it >> 1) & 1u;
      float2* et = ebuf + e * ET;
      mbar_wait(&efull[e], ph);
      if (lane == 0 && it >= 1) { bulk_wait_read<0>(); mbar_arrive(&efree[e ^ 1]); }   // my store of pair it-1 has read its buffer
      if (act_b) { rows_v3_stage_b<S>(et, lane); fence_async_smem(); }
      __syncwarp();
      if (lane == 0) {
        const int g = it0 + it;
        const int slot = v3_slot(ck, B, r, ck.dd0 + g / B, g % B);
        bulk_s2g(scratch + ((long long)slot * N1 + row0) * NP, et, ebytes);
        bulk_commit();
      }
    }
    if (lane == 0) bulk_wait_all<0>();                         // shared memory must outlive the last stores
  }
}

// =========================================================================== forward rows kernel, two roles
// The length-N2 row transforms of the forward pass (capture spectra X[d][b] and replica spectra C[r]) in the
// structure of k_corr_rows_v6: the rows of a 2-row tile arrive by one bulk copy per transform, warp 0 runs the
// forward radix-RA butterflies over the stride-RB digit (the Good-Thomas input permutation fpos2 is applied as it
// reads: its RA source offsets are per-lane constants), warp 1 the radix-RB butterflies, and every stage-B lane
// writes its finished group of RB contiguous outputs back in place with its own 256-byte bulk store. Replaces the
// register-loading k_fwd_rows_s (48 % long-scoreboard stalls, profiles/r04_fwd_ncu_summary.txt) for 480 = 15 x 32.
// grid = (row tiles, splits of the transform list); X: [nt][N] in place.
template <class S> __host__ __device__ constexpr size_t fwd_rows_v6_smem() { return rows_v6_smem<S>(); }
template <class S, int MINCTAS>
__global__ void __launch_bounds__(64, MINCTAS)
k_fwd_rows_v6(DevPlan pl, float2* __restrict__ X, int nt) {
  GNSSACQ_DYN_SMEM(float2, smem);
  static_assert(S::NS == 2 && S::kPfa, "two coprime stages");
  constexpr int T = 2;
  constexpr int N2 = S::F, RA = S::radix(0), RB = S::radix(1), PB = v3_pitch(RB), NP = RA * PB;
  static_assert(RB == 32 && T * RA <= 32, "one lane per stage-A column, stage B within one warp");
  constexpr int XT = T * N2, ET = T * NP;
  float2* xbuf = smem;
  float2* ebuf = smem + 2 * XT;
  unsigned long long* xfull = reinterpret_cast<unsigned long long*>(ebuf + 2 * ET);
  unsigned long long* efull = xfull + 2;
  unsigned long long* efree = xfull + 4;
  int* src_of_pos = reinterpret_cast<int*>(ebuf);               // inverse of fpos2 (tile position -> sample index within the row), built in the
                                                                // first exchange buffer and read into registers before that buffer is used
  const int N = pl.N, N1 = pl.N1;
  const int row0 = blockIdx.x * T;
  const int nrows = imin(T, N1 - row0);
  const int per = (nt + gridDim.y - 1) / gridDim.y;
  const int it0 = blockIdx.y * per, nit = imin(per, nt - it0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned xbytes = (unsigned)nrows * N2 * sizeof(float2);

  auto rows_of = [&](int it) -> float2* { return X + (long long)(it0 + it) * N + (long long)row0 * N2; };
  auto issue = [&](int it) {                                   // rows of transform it0 + it -> ring slot it & 1
    mbar_arrive_expect(&xfull[it & 1], xbytes);
    bulk_g2s(xbuf + (it & 1) * XT, rows_of(it), xbytes, &xfull[it & 1]);
  };
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&xfull[s], 1); mbar_init(&efull[s], 32); mbar_init(&efree[s], 1); }
    mbar_fence_init();
    for (int it = 0; it < 2 && it < nit; ++it) issue(it);
  }
  for (int e = tid; e < N2; e += 64) src_of_pos[__ldg(&pl.fpos2[e])] = e;
  __syncthreads();                                             // barriers and the inverse map visible to both warps

  if (warp == 0) {
    int src[RA];                                               // stage A of column b = lane reads tile positions a * RB + b
#pragma unroll
    for (int a = 0; a < RA; ++a) src[a] = src_of_pos[a * RB + lane];
    __syncwarp();                                              // every lane holds its offsets before the first stage-A store reuses the buffer
    for (int it = 0; it < nit; ++it) {
      const int e = it & 1;
      const unsigned ph = (unsigned)(it >> 1) & 1u;
      float2* et = ebuf + e * ET;
      mbar_wait(&xfull[e], ph);
      if (it >= 2) mbar_wait(&efree[e], ph ^ 1u);              // the stores of transform it-2 have drained this buffer
      for (int ra = 0; ra < nrows; ++ra) {
        const float2* xp = xbuf + e * XT + ra * N2;
        float2 y[RA];
#pragma unroll
        for (int a = 0; a < RA; ++a) y[a] = xp[src[a]];
        Dft<RA>::run(y);
        float2* ep = et + ra * NP + lane;
#pragma unroll
        for (int a = 0; a < RA; ++a) ep[a * PB] = y[a];
      }
      __syncwarp();                                            // every lane is done with ring slot e
      if (lane == 0 && it + 2 < nit) { fence_async_smem(); issue(it + 2); }
      mbar_arrive(&efull[e]);
    }
  } else {
    const bool act_b = lane < nrows * RA;                      // stage-B butterfly: group lane = row * RA + a'
    const int brow = lane / RA, ba = lane - brow * RA;
    for (int it = 0; it < nit; ++it) {
      const int e = it & 1;
      const unsigned ph = (unsigned)(it >> 1) & 1u;
      float2* et = ebuf + e * ET;
      mbar_wait(&efull[e], ph);
      if (it >= 1) {
        if (act_b) bulk_wait_read<0>();                        // my store of transform it-1 has read its group
        __syncwarp();
        if (lane == 0) mbar_arrive(&efree[e ^ 1]);
      }
      if (act_b) {
        float4* p4 = reinterpret_cast<float4*>(et + lane * PB);
        float2 v[RB];
#pragma unroll
        for (int q = 0; q < RB / 2; ++q) { const float4 t = p4[q]; v[2 * q] = make_float2(t.x, t.y); v[2 * q + 1] = make_float2(t.z, t.w); }
        Dft<RB>::run(v);
#pragma unroll
        for (int q = 0; q < RB / 2; ++q) p4[q] = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
        fence_async_smem();                                    // these writes are read by my bulk store
        bulk_s2g(rows_of(it) + brow * N2 + ba * RB, et + lane * PB, (unsigned)(RB * sizeof(float2)));
        bulk_commit();
      }
    }
    if (act_b) bulk_wait_all<0>();                             // shared memory must outlive the last stores
  }
}

// =========================================================================== cols kernel
// ring slots start on 128-byte boundaries (tensor-map copies need it)
template <class S, int CW> __host__ __device__ constexpr int cols_v3_slot() { return (S::F * CW + 15) / 16 * 16; }   // float2 per slot
template <class S, bool MULTI, int CW> __host__ __device__ constexpr size_t cols_v3_smem() {
  return (size_t)2 * cols_v3_slot<S, CW>() * sizeof(float2) + (MULTI ? (size_t)S::F * CW * sizeof(float) : 0) + 16;
}

// Last inverse stage (odd prime radix R0 at stride m0, one thread per butterfly) fused with |.|, the
// non-coherent sum and the peak search. Written for the issue-slot budget ncu showed (profiles/
// README.md, r03b): no re/im swaps (inverse output k is re + i*im of the symmetric sums, output
// R0-k is re - i*im), no per-thread column mask in the hot path, dump and block accumulation as
// template flags, and a peak search whose common path is one max and one compare per batch of
// outputs: a candidate is looked at only if it reaches `floor_` = the best eligible value this
// thread, its warp or — through `unit_hint` — any finished tile of the same (replica, Doppler) unit
// has seen. Any such value is a lower bound of the unit's maximum, so no candidate for the unit's
// (maximum, lowest lag) is ever skipped; tiles that cannot hold it simply report nothing.
// Inputs of the radix-R0 butterfly at tile offset i (column tc): x0, the symmetric / antisymmetric
// pairs a[j] = x[j] + x[R0-j], bq[j] = x[j] - x[R0-j], and output 0 = the plain sum.
template <class S, int CW>
__device__ __forceinline__ void cols_v3_inputs(const float2* tile, int i, int tc, float2& x0, float2* a, float2* bq, float2& s0) {
  constexpr int R0 = S::radix(0), m0 = S::stride(0), H = (R0 - 1) / 2;
  const float2* p = tile + i * CW + tc;
  x0 = p[0];
  s0 = x0;
  static_for<1, H + 1>([&](auto J) {
    constexpr int j = decltype(J)::value;
    const float2 u = p[j * m0 * CW], w = p[(R0 - j) * m0 * CW];
    a[j] = cadd(u, w);
    if constexpr (kRader31On && R0 == 31) {
      bq[j] = r31_flip_b(j) ? csub(w, u) : csub(u, w);         // signed for the sine convolution; output 0 comes from its column sums
    } else {
      bq[j] = csub(u, w);
      s0 = cadd(s0, a[j]);
    }
  });
}

// Outputs of that butterfly, straight into |.|, the non-coherent sum and the peak search.
// QREG (with MULTI): the non-coherent sums of this thread's R0 outputs live in the register array qacc[R0] across
// the blocks of a task instead of in shared memory (one FADD per output and block instead of LDS + FADD + STS);
// the slot of an output is fixed at compile time: 0 for output 0, then in the order the outputs are produced.
template <class S, bool MULTI, bool DUMP, int CW, bool QREG = false>
__device__ __forceinline__ void cols_v3_outputs(const float2 x0, const float2* a, const float2* bq, const float2 s0, int i, float* qs, int tc,
                                               const DevPlan& pl, int lagc, int b, bool last, int n_lags, float scale, float* qd, float hint,
                                               float& best, int& bestlag, float& sum, float* qacc = nullptr) {
  constexpr int R0 = S::radix(0), m0 = S::stride(0), H = (R0 - 1) / 2;
  const int N2 = pl.N2, Nfull = pl.N;
  auto lag_of = [&](int n1i, int q) -> int {
    int n1;
    if constexpr (S::kPfa) { n1 = n1i + q * (S::F / R0); n1 = n1 >= S::F ? n1 - S::F : n1; }
    else n1 = n1i + q * m0;
    const int l = n1 * N2 + lagc;
    return l >= Nfull ? l - Nfull : l;                          // wraps only for coprime splits
  };
  float* qp = qs + i * CW + tc;
  int n1i = i;
  if constexpr (S::kPfa) n1i = __ldg(&pl.n1_of_pos[i]);
  float floor_ = fmaxf(fmaxf(best, hint), 0.f);
  // outputs v[t] with digits Q...: magnitudes, accumulation over blocks, sum, peak candidates
  auto sink = [&](auto digits, auto slot0, const float2* v) {
    constexpr auto dg = seq_array(decltype(digits){});
    constexpr int NB = (int)dg.n, q0 = decltype(slot0)::value;
    float acc[NB];
#pragma unroll
    for (int t = 0; t < NB; ++t) acc[t] = sqrt_fast(fmaf(v[t].x, v[t].x, v[t].y * v[t].y));
    if constexpr (MULTI && QREG) {
#pragma unroll
      for (int t = 0; t < NB; ++t) {
        if (b > 0) acc[t] += qacc[q0 + t];
        qacc[q0 + t] = acc[t];
      }
      if (!last) return;
    } else if constexpr (MULTI) {
#pragma unroll
      for (int t = 0; t < NB; ++t) {
        float* q1 = qp + dg.v[t] * m0 * CW;
        if (b > 0) acc[t] += *q1;
        if (!last) *q1 = acc[t];
      }
      if (!last) return;
    }
    float m = acc[0];
#pragma unroll
    for (int t = 0; t < NB; ++t) sum += acc[t];
#pragma unroll
    for (int t = 1; t < NB; ++t) m = fmaxf(m, acc[t]);
    if (m >= floor_) {
      // rare once the floor is warm. The digit goes through an opaque move so that the lag arithmetic
      // stays inside the branch (ptxas otherwise hoists ~100 integer instructions per butterfly into
      // the common path).
#pragma unroll
      for (int t = 0; t < NB; ++t) {
        if (acc[t] >= floor_) {
          int q = dg.v[t];
#if defined(__CUDA_ARCH__)
          asm volatile("" : "+r"(q));
#endif
          const int lag = lag_of(n1i, q);
          if (lag < n_lags && (acc[t] > best || (acc[t] == best && lag < bestlag))) { best = acc[t]; bestlag = lag; }
        }
      }
      floor_ = fmaxf(floor_, best);
#if defined(__CUDA_ARCH__)
      // share the new floor with the lanes that came along (a lower bound of the unit maximum)
      floor_ = __uint_as_float(__reduce_max_sync(__activemask(), __float_as_uint(floor_)));
#endif
    }
    if constexpr (DUMP) {
#pragma unroll
      for (int t = 0; t < NB; ++t) qd[lag_of(n1i, dg.v[t])] = acc[t] * scale;
    }
  };
  if constexpr (kRader31On && R0 == 31) {
    // two 15-point convolutions (fft_core.cuh); block 2 arrives negated, |.| does not see it
    rader31_outputs(x0, a, bq, [&](const float2 dc) { sink(std::integer_sequence<int, 0>{}, std::integral_constant<int, 0>{}, &dc); },
                    [&](auto N3, const float2* re, const float2* im) {
                      float2 v[10];
#pragma unroll
                      for (int t = 0; t < 5; ++t) {
                        v[2 * t] = make_float2(re[t].x - im[t].y, re[t].y + im[t].x);
                        v[2 * t + 1] = make_float2(re[t].x + im[t].y, re[t].y - im[t].x);
                      }
                      sink(r31_block_digits<decltype(N3)::value>(std::make_integer_sequence<int, 10>{}),
                           std::integral_constant<int, 1 + 10 * decltype(N3)::value>{}, v);
                    });
  } else {
    sink(std::integer_sequence<int, 0>{}, std::integral_constant<int, 0>{}, &s0);
    prime_outputs_batched<R0, 1, H>(x0, a, bq, [&](auto K0c, auto NKc, const float2* re, const float2* im) {
      constexpr int k0 = decltype(K0c)::value, nk = decltype(NKc)::value;
      float2 v[2 * nk];
#pragma unroll
      for (int t = 0; t < nk; ++t) {
        v[2 * t] = make_float2(re[t].x - im[t].y, re[t].y + im[t].x);          // inverse output k      = re + i*im
        v[2 * t + 1] = make_float2(re[t].x + im[t].y, re[t].y - im[t].x);      // inverse output R0 - k = re - i*im
      }
      sink(prime_pair_digits<R0, k0>(std::make_integer_sequence<int, 2 * nk>{}), std::integral_constant<int, 2 * k0 - 1>{}, v);
    });
  }
}

template <class S, bool MULTI, bool DUMP, int CW, int THREADS, bool QREG = false>
__device__ __forceinline__ void cols_v3_last(const float2* tile, float* qs, const DevPlan& pl, int lagc, int b, bool last,
                                            int n_lags, float scale, float* qd, float hint,
                                            float& best, int& bestlag, float& sum, float* qacc = nullptr) {
  constexpr int R0 = S::radix(0), m0 = S::stride(0), H = (R0 - 1) / 2;
  static_assert(R0 % 2 == 1 && R0 >= 7, "prime radix fused with the epilogue");
  const int tc = threadIdx.x & (CW - 1), tb = threadIdx.x / CW;
  constexpr int nb = THREADS / CW;
  static_assert(!QREG || nb >= m0, "register accumulators: one butterfly per thread");
  for (int i = tb; i < m0; i += nb) {
    float2 a[H + 1], bq[H + 1], x0, s0;
    cols_v3_inputs<S, CW>(tile, i, tc, x0, a, bq, s0);
    cols_v3_outputs<S, MULTI, DUMP, CW, QREG>(x0, a, bq, s0, i, qs, tc, pl, lagc, b, last, n_lags, scale, qd, hint, best, bestlag, sum, qacc);
  }
}

// First inverse stage of a columns tile [N1][CW]: radix(1) at unit stride, in place.
template <class S, int CW, int THREADS>
__device__ __forceinline__ void cols_v3_first(float2* tile) {
  constexpr int R = S::radix(1), nbf = S::F / R, nb = THREADS / CW;
  const int tc = threadIdx.x & (CW - 1), tb = threadIdx.x / CW;
#pragma unroll 1
  for (int bf = tb; bf < nbf; bf += nb) {
    float2* p = tile + bf * R * CW + tc;
    float2 v[R];
#pragma unroll
    for (int q = 0; q < R; ++q) v[q] = p[q * CW];
    inv_dft<R>(v);
#pragma unroll
    for (int q = 0; q < R; ++q) p[q * CW] = v[q];
  }
}

// Persistent: task t = blockIdx.x + k * gridDim.x = (unit ul of the chunk, column tile ct), B items each.
// map: 8-byte elements, dims (NP, F1, F2 * slots), box (CW, F1, F2) with N1 = F1 * F2; zmul = F2.
// pl.col_lag must point at the padded column table (NP + slack entries, -1 = pad column).
// unit_hint[r*D + d]: float bits of the best eligible value any finished tile of that unit has
// reported (zeroed by the host before the search); read at the start of a task, raised at its end.
template <class S, bool MULTI, bool DUMP, int CW, int THREADS, int MINCTAS, bool CHORE_LAST = true, bool QREG = false>
__global__ void __launch_bounds__(THREADS, MINCTAS)
k_corr_cols_v3(DevPlan pl, const GNSSACQ_GRID_CONSTANT TensorMap map, int zmul, const int* __restrict__ tile_col0,
               ChunkV3 ck, int B, int D, int d0, int n_lags, float scale, int ntiles,
               Part* __restrict__ parts, float* __restrict__ q_dump, unsigned* __restrict__ unit_hint) {
  GNSSACQ_DYN_SMEM(float2, smem);
  static_assert(S::NS == 2, "two-stage columns schedule");
  static_assert(CW == 8 || CW == 16, "tile width");
  constexpr int N1 = S::F, TILE = N1 * CW, SLOT = cols_v3_slot<S, CW>();
  float* qs = reinterpret_cast<float*>(smem + 2 * SLOT);          // unused with QREG (the launch then passes the size without it)
  unsigned long long* full = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(smem) + cols_v3_smem<S, MULTI && !QREG, CW>() - 16);
  float qacc[QREG ? S::radix(0) : 1];
  const int N = pl.N;
  const int tid = threadIdx.x, tc = tid & (CW - 1);
  static_assert(THREADS % 32 == 0, "whole warps");
  constexpr int NW = THREADS / 32;
  // The per-tile chores (refilling the ring, folding the previous tile's parts) belong to the first lane of the LAST
  // warp: the first stage has nbf * CW butterflies for THREADS threads, and it is the last warp that has a round less.
  constexpr int kChore = CHORE_LAST ? THREADS - 32 : 0;            // 0: A/B variant
  // per-warp results of a finished tile; thread 0 folds them behind the next block barrier (two sets: the
  // warps of the next tile write the other one)
  __shared__ unsigned long long s_key[2][NW];
  __shared__ float s_sum[2][NW];
  // Tasks are tile-major (task = ct * units + ul, ul = ur * G + ud): the tiles of one unit are spread over the
  // waves of the persistent grid, so all but the first carry a peak-search floor from the unit's earlier
  // tiles. A CTA's tasks are gridDim.x apart; (ct, ur, ud) advance by carries instead of divisions.
  const int nunits = ck.Rc * ck.G;
  const int step_ct = (int)gridDim.x / nunits, step_ul = (int)gridDim.x - step_ct * nunits;
  const int step_ur = step_ul / ck.G, step_ud = step_ul - step_ur * ck.G;
  int ct = (int)blockIdx.x / nunits, ur, ud;
  { const int ul = (int)blockIdx.x - ct * nunits; ur = ul / ck.G; ud = ul - ur * ck.G; }
  auto issue = [&](int ct_, int ul, int b, int slot) {         // thread 0
    mbar_arrive_expect(&full[slot], (unsigned)(TILE * sizeof(float2)));
    tma_load_3d(smem + slot * SLOT, &map, __ldg(&tile_col0[ct_]), 0, (ul * B + b) * zmul, &full[slot]);
  };
  int b = 0;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
    tma_prefetch_map(&map);
    if (ct < ntiles) issue(ct, ur * ck.G + ud, 0, 0);
  }
  __syncthreads();
  float best = -1.f, sum = 0.f, hint = 0.f;
  int bestlag = 0x7fffffff, lagc = -1;
  long long pend_unit = -1;                                    // chore thread: finished tile whose part is not written yet
  int pend_ct = 0, pend_set = 0;
  auto flush = [&]() {                                         // chore thread, behind a block barrier
    unsigned long long key = s_key[pend_set][0];
    float sm = s_sum[pend_set][0];
#pragma unroll
    for (int w = 1; w < NW; ++w) { const unsigned long long k2 = s_key[pend_set][w]; key = k2 > key ? k2 : key; sm += s_sum[pend_set][w]; }
    Part p; p.key = 0ull; p.sum = sm; p.pad = 0.f;
    if (key != 0ull) {
      // the hint must be bit-exactly a value that occurred, so the search runs on unscaled values; scaling by
      // 1/N is monotonic: order and ties of the keys are those of the scaled values
      const unsigned vb = (unsigned)(key >> 32);
      p.key = ((unsigned long long)__float_as_uint(__uint_as_float(vb) * scale) << 32) | (key & 0xffffffffull);
      atomicMax(&unit_hint[pend_unit], vb);
    }
    parts[pend_unit * ntiles + pend_ct] = p;
    pend_unit = -1;
  };
  for (unsigned seq = 0; ct < ntiles; ++seq) {
    int nct = ct, nur = ur, nud = ud, nblk = b + 1;
    if (nblk == B) {
      nblk = 0;
      nct += step_ct; nur += step_ur; nud += step_ud;
      if (nud >= ck.G) { nud -= ck.G; ++nur; }
      if (nur >= ck.Rc) { nur -= ck.Rc; ++nct; }
    }
    __syncthreads();                                           // every thread is done with the previous item: its slot is free
    if (tid == kChore) {
      if (nct < ntiles) { fence_async_smem(); issue(nct, nur * ck.G + nud, nblk, (seq + 1) & 1); }
      if (pend_unit >= 0) flush();
    }
    const long long unit = (long long)(ck.r0 + ur) * D + d0 + ck.dd0 + ud;
    const bool last = (b + 1 == B);
    if (b == 0) {
      best = -1.f; sum = 0.f; bestlag = 0x7fffffff;
      lagc = __ldg(&pl.col_lag[__ldg(&tile_col0[ct]) + tc]);     // -1: pad column, takes no part in the epilogue
      hint = __uint_as_float(__ldcg(&unit_hint[unit]));         // unscaled, exactly a value some tile has seen (may be stale: any lower bound will do)
    }
    float* qd = DUMP ? q_dump + unit * N : nullptr;
    float2* tile = smem + (seq & 1) * SLOT;
    mbar_wait(&full[seq & 1], (seq >> 1) & 1u);
    cols_v3_first<S, CW, THREADS>(tile);
    __syncthreads();
    if (lagc >= 0) cols_v3_last<S, MULTI, DUMP, CW, THREADS, QREG>(tile, qs, pl, lagc, b, last, n_lags, scale, qd, hint, best, bestlag, sum, qacc);
    if (last) {
      // per warp: the sum always, the key only if some lane holds a candidate (rare once the floor is warm)
      unsigned long long key = bestlag != 0x7fffffff ? pack_key(best, bestlag) : 0ull;
      float sm = sum * scale;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
      if (__any_sync(0xffffffffu, key != 0ull)) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o);
          key = k2 > key ? k2 : key;
        }
      }
      if ((tid & 31) == 0) { s_key[seq & 1][tid >> 5] = key; s_sum[seq & 1][tid >> 5] = sm; }
      if (tid == kChore) { pend_unit = unit; pend_ct = ct; pend_set = (int)(seq & 1); }
    }
    ct = nct; ur = nur; ud = nud; b = nblk;
  }
  __syncthreads();
  if (tid == kChore && pend_unit >= 0) flush();
}

// =========================================================================== cols kernel, one tile slot
// The radix-31 stage reads its 31 inputs at once and then computes ~500 instructions from registers, so the tile
// slot is free long before the tile is finished: the next tensor-map copy is issued as soon as every thread has
// arrived at the `empty` mbarrier with its inputs in registers, into the SAME slot. Half the shared memory of
// k_corr_cols_v3, so registers alone decide how many CTAs an SM holds. Same task order, same per-warp parts folded
// by thread 0 behind the block barrier of the next tile, same arithmetic: bit-identical results.
// Two things this kernel has to get right (profiles/README.md, r04): the butterfly's inputs must not cross a
// control-flow merge between the loads and the epilogue (one region, or ptxas spends ~40 registers on re-pairing
// them), and lanes without a butterfly must arrive BEFORE the others enter that region (a diverged warp runs its
// active lanes through the whole epilogue first).
template <class S, bool MULTI, int CW> __host__ __device__ constexpr size_t cols_v5_smem() {
  return (size_t)cols_v3_slot<S, CW>() * sizeof(float2) + (MULTI ? (size_t)S::F * CW * sizeof(float) : 0) + 16;
}
template <class S, bool MULTI, bool DUMP, int CW, int THREADS, int MINCTAS>
__global__ void __launch_bounds__(THREADS, MINCTAS)
k_corr_cols_v5(DevPlan pl, const GNSSACQ_GRID_CONSTANT TensorMap map, int zmul, const int* __restrict__ tile_col0,
               ChunkV3 ck, int B, int D, int d0, int n_lags, float scale, int ntiles,
               Part* __restrict__ parts, float* __restrict__ q_dump, unsigned* __restrict__ unit_hint) {
  GNSSACQ_DYN_SMEM(float2, smem);
  static_assert(S::NS == 2, "two-stage columns schedule");
  static_assert(CW == 8 || CW == 16, "tile width");
  constexpr int N1 = S::F, TILE = N1 * CW, SLOT = cols_v3_slot<S, CW>();
  constexpr int R0 = S::radix(0), m0 = S::stride(0), H = (R0 - 1) / 2;
  static_assert(THREADS / CW >= m0 && 32 / CW <= m0, "one radix-R0 butterfly per thread; warp 0 holds butterflies only");
  float* qs = reinterpret_cast<float*>(smem + SLOT);
  // full: the copy engine has filled the slot; empty: every thread has its radix-R0 inputs of the tile in registers
  unsigned long long* full = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(smem) + cols_v5_smem<S, MULTI, CW>() - 16);
  unsigned long long* empty = full + 1;
  const int N = pl.N;
  const int tid = threadIdx.x, tc = tid & (CW - 1), tb = tid / CW;
  static_assert(THREADS % 32 == 0, "whole warps");
  constexpr int NW = THREADS / 32;
  // per-warp results of a finished tile; thread 0 folds them behind the next block barrier (two sets: the
  // warps of the next tile write the other one)
  __shared__ unsigned long long s_key[2][NW];
  __shared__ float s_sum[2][NW];
  // Tasks are tile-major (task = ct * units + ul, ul = ur * G + ud): the tiles of one unit are spread over the
  // waves of the persistent grid, so all but the first carry a peak-search floor from the unit's earlier
  // tiles. A CTA's tasks are gridDim.x apart; (ct, ur, ud) advance by carries instead of divisions.
  struct Pos { int ct, ur, ud, b; };
  const int nunits = ck.Rc * ck.G;
  const int step_ct = (int)gridDim.x / nunits, step_ul = (int)gridDim.x - step_ct * nunits;
  const int step_ur = step_ul / ck.G, step_ud = step_ul - step_ur * ck.G;
  auto advance = [&](Pos p) -> Pos {                            // next item of this CTA: next block, else next task
    if (++p.b == B) {
      p.b = 0;
      p.ct += step_ct; p.ur += step_ur; p.ud += step_ud;
      if (p.ud >= ck.G) { p.ud -= ck.G; ++p.ur; }
      if (p.ur >= ck.Rc) { p.ur -= ck.Rc; ++p.ct; }
    }
    return p;
  };
  auto issue = [&](const Pos& p) {
    mbar_arrive_expect(full, (unsigned)(TILE * sizeof(float2)));
    tma_load_3d(smem, &map, __ldg(&tile_col0[p.ct]), 0, ((p.ur * ck.G + p.ud) * B + p.b) * zmul, full);
  };
  Pos cur;
  cur.ct = (int)blockIdx.x / nunits; cur.b = 0;
  { const int ul = (int)blockIdx.x - cur.ct * nunits; cur.ur = ul / ck.G; cur.ud = ul - cur.ur * ck.G; }
  Pos nx = advance(cur);
  if (tid == 0) {
    mbar_init(full, 1);
    mbar_init(empty, THREADS);
    mbar_fence_init();
    tma_prefetch_map(&map);
    if (cur.ct < ntiles) issue(cur);
  }
  __syncthreads();
  float best = -1.f, sum = 0.f, hint = 0.f;
  int bestlag = 0x7fffffff, lagc = -1;
  long long pend_unit = -1;                                    // thread 0: finished tile whose part is not written yet
  int pend_ct = 0, pend_set = 0;
  auto flush = [&]() {                                         // thread 0, behind a block barrier
    unsigned long long key = s_key[pend_set][0];
    float sm = s_sum[pend_set][0];
#pragma unroll
    for (int w = 1; w < NW; ++w) { const unsigned long long k2 = s_key[pend_set][w]; key = k2 > key ? k2 : key; sm += s_sum[pend_set][w]; }
    Part p; p.key = 0ull; p.sum = sm; p.pad = 0.f;
    if (key != 0ull) {
      // the hint must be bit-exactly a value that occurred, so the search runs on unscaled values; scaling by
      // 1/N is monotonic: order and ties of the keys are those of the scaled values
      const unsigned vb = (unsigned)(key >> 32);
      p.key = ((unsigned long long)__float_as_uint(__uint_as_float(vb) * scale) << 32) | (key & 0xffffffffull);
      atomicMax(&unit_hint[pend_unit], vb);
    }
    parts[pend_unit * ntiles + pend_ct] = p;
    pend_unit = -1;
  };
  for (unsigned seq = 0; cur.ct < ntiles; ++seq) {
    const int sl = (int)(seq & 1);                             // set of per-warp results
    const unsigned ph = seq & 1u;
    const int b = cur.b;
    const long long unit = (long long)(ck.r0 + cur.ur) * D + d0 + ck.dd0 + cur.ud;
    const bool last = (b + 1 == B);
    if (b == 0) {
      best = -1.f; sum = 0.f; bestlag = 0x7fffffff;
      lagc = __ldg(&pl.col_lag[__ldg(&tile_col0[cur.ct]) + tc]);  // -1: pad column, takes no part in the epilogue
      hint = __uint_as_float(__ldcg(&unit_hint[unit]));         // unscaled, exactly a value some tile has seen (may be stale: any lower bound will do)
    }
    float* qd = DUMP ? q_dump + unit * N : nullptr;
    float2* tile = smem;
    mbar_wait(full, ph);
    cols_v3_first<S, CW, THREADS>(tile);
    __syncthreads();                                           // the only block barrier of a tile
    if (tid == 0 && pend_unit >= 0) flush();                   // every warp has left its results of the previous tile
    // pad columns and spare butterfly slots only keep the barriers. They arrive BEFORE the others' region: a warp
    // that diverges runs its active lanes through the whole epilogue first, and the refill below would wait for that.
    const bool act = lagc >= 0 && tb < m0;
    const unsigned actmask = __ballot_sync(0xffffffffu, act);
    // the refill is the duty of the first active lane of warp 0 (warp 0 holds butterflies only, tb < m0), of lane 0 if a
    // tile has no real column at all: never of a lane that waits while lanes of its own warp still have to arrive
    const bool refiller = tid < 32 && tid == (actmask ? __ffs((int)actmask) - 1 : 0);
    auto refill = [&]() {                                      // refill the slot with the next item
      if (nx.ct < ntiles) {
        mbar_wait(empty, ph);
        fence_async_smem();
        issue(nx);
      }
    };
    if (!act) {
      mbar_arrive(empty);
      if (refiller) refill();
    }
    if (act) {
      // one region from the loads to the epilogue: the inputs stay in the register pairs the packed instructions need
      float2 a[H + 1], bq[H + 1], x0, s0;
      cols_v3_inputs<S, CW>(tile, tb, tc, x0, a, bq, s0);
      mbar_arrive(empty);                                 // my inputs are in registers
      if (refiller) refill();
      cols_v3_outputs<S, MULTI, DUMP, CW>(x0, a, bq, s0, tb, qs, tc, pl, lagc, b, last, n_lags, scale, qd, hint, best, bestlag, sum);
    }
    if (last) {
      // per warp: the sum always, the key only if some lane holds a candidate (rare once the floor is warm)
      unsigned long long key = bestlag != 0x7fffffff ? pack_key(best, bestlag) : 0ull;
      float sm = sum * scale;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
      if (__any_sync(0xffffffffu, key != 0ull)) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o);
          key = k2 > key ? k2 : key;
        }
      }
      if ((tid & 31) == 0) { s_key[sl][tid >> 5] = key; s_sum[sl][tid >> 5] = sm; }
      if (tid == 0) { pend_unit = unit; pend_ct = cur.ct; pend_set = sl; }
    }
    cur = nx; nx = advance(nx);
  }
  __syncthreads();
  if (tid == 0 && pend_unit >= 0) flush();
}

}  // namespace acq
