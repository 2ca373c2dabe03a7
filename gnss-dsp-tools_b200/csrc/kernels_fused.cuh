// One persistent kernel for the whole correlate stage of a Doppler chunk (coprime plans with
// two-stage schedules, as kernels_v3.cuh): every CTA draws tasks from one ordered ticket stream,
//
//     R(0), R(1), C(0), R(2), C(1), ..., R(n-1), C(n-2), C(n-1)
//
// where R(k) are the rows tasks of unit group k (replica-spectrum multiply + length-N2 inverse
// transforms into the scratch set k mod NSETS) and C(k) its columns tasks (length-N1 inverse
// transforms + |.| + non-coherent sum + peak search). A group is Rc replicas x G Doppler bins x B
// blocks, small enough that NSETS scratch sets stay L2-resident, so the intermediate spectra of
// the whole search never travel to HBM — with two kernels per chunk that held only for chunks too
// small to fill the GPU (profiles/README.md r03d: 8 GB of DRAM traffic per 2 ms step). There are no
// launch boundaries, hence no per-launch tails, and rows and columns tasks of neighbouring groups
// share every SM.
//
// Dependencies are two counters per group in global memory: C(k) waits until all R(k) tasks have
// published their tiles (bulk stores complete, then a release increment), R(k) waits until all
// C(k - NSETS) tasks have consumed the set it is about to overwrite. Every task a ticket can wait
// for was handed out earlier, hence to a CTA that is running (tickets are drawn only by resident
// CTAs), so waits are bounded whatever the grid size; a spin limit turns a protocol bug into an
// error flag instead of a hung GPU. Task bodies are those of kernels_v3.cuh.
#pragma once
#include "kernels_v3.cuh"

namespace acq {

template <class SR, int T> __host__ __device__ constexpr int fused_rows_floats2() {
  return T * SR::F + 2 * T * SR::radix(0) * v3_pitch(SR::radix(1));
}
template <class SR, class SC, int T, int CW, bool MULTI> __host__ __device__ constexpr size_t fused_smem() {
  const size_t rows = (size_t)fused_rows_floats2<SR, T>() * sizeof(float2);
  const size_t cols = (size_t)2 * cols_v3_slot<SC, CW>() * sizeof(float2) + (MULTI ? (size_t)SC::F * CW * sizeof(float) : 0);
  return (rows > cols ? rows : cols) + 64;           // + mbarriers (xfull, cfull[2]) and the task broadcast
}

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void spin_pause() { __nanosleep(64); }
#elif defined(GNSSACQ_EMU_BUILD)
inline int ld_acquire(const int* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void red_release_add(int* p, int v) { __atomic_fetch_add(p, v, __ATOMIC_RELEASE); }
inline void fence_async_all() {}
inline void spin_pause() { std::this_thread::yield(); }
#else      // host pass of nvcc: never executed
__device__ __forceinline__ int ld_acquire(const int* p) { return *p; }
__device__ __forceinline__ void red_release_add(int*, int) {}
__device__ __forceinline__ void fence_async_all() {}
__device__ __forceinline__ void spin_pause() {}
#endif

// Thread 0: wait until *ctr >= target. Returns false (and raises the error flag) after ~seconds.
__device__ __forceinline__ bool fused_wait(const int* ctr, int target, int* error) {
  for (long long spins = 0; ld_acquire(ctr) < target; ++spins) {
    if (spins > (1ll << 24)) { atomicAdd(error, 1); return false; }
    spin_pause();
  }
  return true;
}

template <class SR, class SC, int T, int CW, bool MULTI, bool DUMP, int THREADS, int MINCTAS>
__global__ void __launch_bounds__(THREADS, MINCTAS)
k_corr_fused(DevPlan pl, const GNSSACQ_GRID_CONSTANT TensorMap map, const int* __restrict__ tile_col0, FusedJob job, FusedSync sy,
             const float2* __restrict__ X, const float2* __restrict__ C, float2* __restrict__ scratch,
             Part* __restrict__ parts, float* __restrict__ q_dump, unsigned* __restrict__ unit_hint) {
  GNSSACQ_DYN_SMEM(float2, smem);
  static_assert(SR::NS == 2 && SR::kPfa && SC::NS == 2, "two-stage schedules");
  constexpr int N2 = SR::F, RA = SR::radix(0), RB = SR::radix(1), PB = v3_pitch(RB), NP = RA * PB;
  constexpr int N1 = SC::F, XT = T * N2, ET = T * NP, TILE = N1 * CW, SLOT = cols_v3_slot<SC, CW>();
  static_assert(RB % 2 == 0 && THREADS >= T * RB && THREADS >= T * RA && THREADS % CW == 0 && THREADS / CW >= SC::stride(0), "thread mapping");
  unsigned char* tail = reinterpret_cast<unsigned char*>(smem) + fused_smem<SR, SC, T, CW, MULTI>() - 64;
  unsigned long long* xfull = reinterpret_cast<unsigned long long*>(tail);            // rows: spectra tile landed
  unsigned long long* cfull = xfull + 1;                                               // cols: tile slot 0 / 1 landed
  int* s_task = reinterpret_cast<int*>(tail + 32);                                     // ticket broadcast
  const int N = pl.N;
  const int tid = threadIdx.x;
  unsigned xuse = 0, cuse0 = 0, cuse1 = 0;            // completed phases of each barrier (uniform across the CTA)

  if (tid == 0) {
    mbar_init(xfull, 1);
    mbar_init(&cfull[0], 1);
    mbar_init(&cfull[1], 1);
    mbar_fence_init();
    tma_prefetch_map(&map);
  }
  __syncthreads();

  const int per_step = job.nR + job.nC, total = job.ng * per_step;
  for (;;) {
    __syncthreads();                                   // everyone is done with the previous task (s_task, shared memory)
    if (tid == 0) s_task[0] = atomicAdd(sy.ticket, 1);
    __syncthreads();
    const int t = s_task[0];
    if (t >= total) break;
    // ---- decode the ticket: R(0) | then steps s = 1..ng of [R(s) (absent in the last), C(s-1)]
    bool is_rows; int k, idx;
    if (t < job.nR) { is_rows = true; k = 0; idx = t; }
    else {
      const int t1 = t - job.nR, s = 1 + t1 / per_step, w = t1 - (s - 1) * per_step;
      if (s < job.ng && w < job.nR) { is_rows = true; k = s; idx = w; }
      else { is_rows = false; k = s - 1; idx = s < job.ng ? w - job.nR : w; }
    }
    const int gd = k / job.ngr, gr = k - gd * job.ngr;
    ChunkV3 ck;
    ck.r0 = gr * job.Rc; ck.Rc = imin(job.Rc, job.R - ck.r0);
    ck.dd0 = gd * job.G; ck.G = imin(job.G, job.dc - ck.dd0);
    const int B = job.B;
    const long long set_base = (long long)(k % job.nsets) * job.slots_per_set;      // first slot of this group's scratch set

    if (is_rows) {
      // =================================================================== rows task (row tile rt, replica rr)
      const int rt = idx % job.nrt, rr = idx / job.nrt;
      const bool live = rr < ck.Rc;
      if (tid == 0 && live && k >= job.nsets) fused_wait(&sy.cols_done[k - job.nsets], job.nC, sy.error);   // the set is free again
      float2* xbuf = smem;
      float2* ebuf = smem + XT;
      const int row0 = rt * T, nrows = imin(T, pl.N1 - row0);
      const int r = ck.r0 + rr;
      const int nit = ck.G * B;
      const unsigned xbytes = (unsigned)nrows * N2 * sizeof(float2), ebytes = (unsigned)nrows * NP * sizeof(float2);
      auto issue = [&](int it) {
        const float2* src = X + ((long long)(ck.dd0 + it / B) * B + it % B) * N + (long long)row0 * N2;
        mbar_arrive_expect(xfull, xbytes);
        bulk_g2s(xbuf, src, xbytes, xfull);
      };
      if (live) {
        if (tid == 0) { fence_async_smem(); issue(0); }
        const int ra = tid / RB, b = tid - ra * RB;
        const bool act_a = tid < T * RB && ra < nrows;
        const bool act_b = tid < nrows * RA;
        float2 c[RA];
        if (act_a) {
          const float2* cp = C + (long long)r * N + (long long)(row0 + ra) * N2 + b;
#pragma unroll
          for (int a = 0; a < RA; ++a) c[a] = __ldg(&cp[a * RB]);
        }
        __syncthreads();                               // thread 0's wait for the free set, seen by all before any store is issued
        for (int it = 0; it < nit; ++it) {
          float2* et = ebuf + (it & 1) * ET;
          mbar_wait(xfull, xuse & 1u);
          ++xuse;
          if (act_a) rows_v3_stage_a<SR>(xbuf, et, c, ra, b);
          __syncthreads();
          if (tid == 0 && it + 1 < nit) { fence_async_smem(); issue(it + 1); }
          if (act_b) { rows_v3_stage_b<SR>(et, tid); fence_async_smem(); }
          if (tid == 0) bulk_wait_read<0>();
          __syncthreads();
          if (tid == 0) {
            const long long slot = set_base + v3_slot(ck, B, r, ck.dd0 + it / B, it % B);
            bulk_s2g(scratch + (slot * pl.N1 + row0) * NP, et, ebytes);
            bulk_commit();
          }
        }
      }
      if (tid == 0) {
        bulk_wait_all<0>();                            // this task's tiles are written ...
        fence_async_all();
        __threadfence();
        red_release_add(&sy.rows_done[k], 1);          // ... and published
      }
    } else {
      // =================================================================== columns task: unit ul, tiles ct0 .. ct0+tpt-1, B items each
      // Several tiles per ticket: the two-slot ring prefetches item i+1 while item i is transformed,
      // and the ticket / dependency / barrier overhead is paid once per tpt tiles.
      const int nunits_max = job.Rc * job.G;
      const int cg = idx / nunits_max, ul_full = idx - cg * nunits_max;
      const int ct0 = cg * job.tpt, nt = imin(job.tpt, job.ntiles - ct0);
      const int rr = ul_full / job.G, dg = ul_full - rr * job.G;
      const bool live = rr < ck.Rc && dg < ck.G && nt > 0;
      if (tid == 0) {
        if (live && !fused_wait(&sy.rows_done[k], job.nR, sy.error)) s_task[1] = 1; else s_task[1] = 0;
        fence_async_all();                             // the tiles were written through the async proxy and will be read through it
      }
      float* qs = reinterpret_cast<float*>(smem + 2 * SLOT);
      const int r = ck.r0 + rr, dd = ck.dd0 + dg;
      const long long unit = (long long)r * job.D + job.d0 + dd;
      const int ul = rr * ck.G + dg;                   // slot numbering of v3_slot
      const int nitems = nt * B;
      auto issue = [&](int item, int slot) {
        const int ct = ct0 + item / B, b = item % B;
        mbar_arrive_expect(&cfull[slot], (unsigned)(TILE * sizeof(float2)));
        tma_load_3d(smem + slot * SLOT, &map, __ldg(&tile_col0[ct]), 0, (int)((set_base + (long long)ul * B + b) * job.zmul), &cfull[slot]);
      };
      __syncthreads();                                 // wait result visible; shared memory of the previous task released
      if (live && s_task[1] == 0) {
        if (tid == 0) { fence_async_smem(); issue(0, 0); }
        const int tc = tid & (CW - 1);
        float* qd = DUMP ? q_dump + unit * N : nullptr;
        float best = -1.f, sum = 0.f, hint = 0.f;
        int bestlag = 0x7fffffff, lagc = -1;
        for (int item = 0; item < nitems; ++item) {
          const int slot = item & 1, ct = ct0 + item / B, b = item % B;
          if (item > 0) __syncthreads();               // everyone is done with the slot the next copy overwrites
          if (tid == 0 && item + 1 < nitems) { fence_async_smem(); issue(item + 1, slot ^ 1); }
          if (b == 0) {
            best = -1.f; sum = 0.f; bestlag = 0x7fffffff;
            lagc = __ldg(&pl.col_lag[__ldg(&tile_col0[ct]) + tc]);
            hint = __uint_as_float(__ldcg(&unit_hint[unit]));
          }
          float2* tile = smem + slot * SLOT;
          if (slot == 0) { mbar_wait(&cfull[0], cuse0 & 1u); ++cuse0; } else { mbar_wait(&cfull[1], cuse1 & 1u); ++cuse1; }
          cols_v3_first<SC, CW, THREADS>(tile);
          __syncthreads();
          if (lagc >= 0)
            cols_v3_last<SC, MULTI, DUMP, CW, THREADS>(tile, qs, pl, lagc, b, b + 1 == B, job.n_lags, job.scale, qd, hint, best, bestlag, sum);
          if (b + 1 == B) {
            unsigned long long key = bestlag != 0x7fffffff ? pack_key(best, bestlag) : 0ull;
            float sm = sum * job.scale;
            block_reduce_part(key, sm);
            if (tid == 0) {
              Part p; p.key = 0ull; p.sum = sm; p.pad = 0.f;
              if (key != 0ull) {
                const unsigned vb = (unsigned)(key >> 32);
                p.key = ((unsigned long long)__float_as_uint(__uint_as_float(vb) * job.scale) << 32) | (key & 0xffffffffull);
                atomicMax(&unit_hint[unit], vb);
              }
              parts[unit * job.ntiles + ct] = p;
            }
          }
        }
      }
      if (tid == 0) {
        __threadfence();
        red_release_add(&sy.cols_done[k], 1);          // the set may be overwritten once all of C(k) got here
      }
    }
  }
}

}  // namespace acq
