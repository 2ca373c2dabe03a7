// Asynchronous-copy primitives of sm_100a: mbarrier transaction barriers, 1-D bulk copies
// (cp.async.bulk, SASS UBLKCP) and tiled tensor-map copies (cp.async.bulk.tensor, SASS
// UTMALDG), plus the host-side tensor-map encoder. The encoder is looked up through the CUDA
// runtime (cudaGetDriverEntryPoint), so libgnssacq.so does not link against libcuda and still
// loads on a machine without a driver (the ABI tests do exactly that).
//
// When GNSSACQ_EMU_BUILD is defined (tests/cuda_emu, a test harness) the same names are host
// functions: copies complete at issue time and an mbarrier is a small record behind a mutex, which
// is enough to run the kernels' index logic and their barrier protocol on the CPU.
#pragma once
#include "cuda_compat.h"

namespace acq {

// Opaque 128-byte tensor map (CUtensorMap on the device; the emulator keeps its own fields in it).
struct alignas(64) TensorMap { unsigned long long opaque[16]; };

#ifndef GNSSACQ_EMU_BUILD
// ------------------------------------------------------------------------------ device
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the initialised barriers visible to the async proxy (the copy engine)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one arrival that also announces `bytes` of copy traffic for the current phase
__device__ __forceinline__ void mbar_arrive_expect(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// plain arrival (no copy traffic announced)
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier over `count` threads of the CTA (a multiple of 32), id 1..15 (0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, contiguous; bytes and both addresses are multiples of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global, contiguous, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups are still *reading* their shared-memory source
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// 3-D tiled tensor-map load: box at coordinates (c0 fastest, c1, c2) -> dense box in shared memory
__device__ __forceinline__ void tma_load_3d(void* dst, const TensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_map(const TensorMap* map) { asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory"); }
// order this thread's earlier generic-proxy accesses of shared memory before later async-proxy ones
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#define GNSSACQ_GRID_CONSTANT __grid_constant__

#else
// ------------------------------------------------------------------------------ host emulation
struct EmuTensorMap { const unsigned char* base; int rank; long long dim[3]; long long stride[3]; int box[3]; int esize; };
static_assert(sizeof(EmuTensorMap) <= sizeof(TensorMap), "emulated tensor map must fit");
void emu_mbar_init(unsigned long long* bar, int count);
void emu_mbar_arrive(unsigned long long* bar, long long tx);
void emu_mbar_complete_tx(unsigned long long* bar, long long bytes);
bool emu_mbar_test(unsigned long long* bar, unsigned parity);
inline void mbar_init(unsigned long long* bar, int count) { emu_mbar_init(bar, count); }
inline void mbar_fence_init() {}
inline void mbar_arrive_expect(unsigned long long* bar, unsigned bytes) { emu_mbar_arrive(bar, bytes); }
inline void mbar_arrive(unsigned long long* bar) { emu_mbar_arrive(bar, 0); }
void emu_named_barrier(int id, int count);
inline void named_bar_sync(int id, int count) { emu_named_barrier(id, count); }
inline void mbar_wait(unsigned long long* bar, unsigned parity) { while (!emu_mbar_test(bar, parity)) std::this_thread::yield(); }
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) { memcpy(dst, src, bytes); emu_mbar_complete_tx(bar, bytes); }
inline void bulk_s2g(void* dst, const void* src, unsigned bytes) { memcpy(dst, src, bytes); }
inline void bulk_commit() {}
template <int N> inline void bulk_wait_read() {}
template <int N> inline void bulk_wait_all() {}
inline void tma_load_3d(void* dst, const TensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
  const EmuTensorMap& m = *reinterpret_cast<const EmuTensorMap*>(map);
  unsigned char* d = static_cast<unsigned char*>(dst);
  long long bytes = 0;
  for (int k = 0; k < m.box[2]; ++k)
    for (int j = 0; j < m.box[1]; ++j)
      for (int i = 0; i < m.box[0]; ++i, d += m.esize, bytes += m.esize) {
        const long long x = c0 + i, y = c1 + j, z = c2 + k;
        if (x < m.dim[0] && y < m.dim[1] && z < m.dim[2]) memcpy(d, m.base + x * m.esize + y * m.stride[1] + z * m.stride[2], m.esize);
        else memset(d, 0, m.esize);                       // out-of-bounds elements read as zero
      }
  emu_mbar_complete_tx(bar, bytes);
}
inline void tma_prefetch_map(const TensorMap*) {}
inline void fence_async_smem() {}
#define GNSSACQ_GRID_CONSTANT
#endif

// Host: tensor map over 8-byte elements, rank 3. dims / box in elements (dim 0 fastest),
// strides in bytes for dims 1 and 2. Returns 0 on success, the driver's error code otherwise.
int encode_tensor_map_3d_u64(TensorMap* out, const void* base, const unsigned long long dims[3],
                             const unsigned long long strides_bytes[2], const unsigned box[3]);

}  // namespace acq
