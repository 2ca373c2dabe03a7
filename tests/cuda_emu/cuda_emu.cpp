// TEST HARNESS ONLY — see cuda_emu.h. Runs one emulated block at a time on a thread pool.
#include "cuda_emu.h"

thread_local EmuIdx threadIdx, blockIdx;
EmuIdx blockDim, gridDim;
std::barrier<>* emu_block_barrier = nullptr;
unsigned char* emu_dyn_smem = nullptr;
thread_local int emu_lane = 0, emu_warp = 0;
EmuWarp* emu_warps = nullptr;

namespace {
struct Pool {
  std::vector<std::thread> th;
  std::mutex mu;
  std::condition_variable cv, cv_done;
  long gen = 0;
  int nactive = 0, remaining = 0;
  EmuIdx bidx{};
  const std::function<void()>* body = nullptr;
  bool quit = false;

  void worker(int id) {
    long seen = 0;
    for (;;) {
      const std::function<void()>* b;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return quit || gen != seen; });
        if (quit) return;
        seen = gen;
        if (id >= nactive) continue;
        b = body;
        blockIdx = bidx;
      }
      threadIdx.x = id % blockDim.x;
      threadIdx.y = (id / blockDim.x) % blockDim.y;
      threadIdx.z = id / (blockDim.x * blockDim.y);
      emu_lane = id & 31;
      emu_warp = id >> 5;
      (*b)();
      {
        std::lock_guard<std::mutex> lk(mu);
        if (--remaining == 0) cv_done.notify_all();
      }
    }
  }
  void ensure(int n) {
    while ((int)th.size() < n) { int id = (int)th.size(); th.emplace_back([this, id] { worker(id); }); }
  }
  void run_block(int n, EmuIdx b, const std::function<void()>& f) {
    ensure(n);
    {
      std::lock_guard<std::mutex> lk(mu);
      nactive = n; remaining = n; bidx = b; body = &f; ++gen;
    }
    cv.notify_all();
    std::unique_lock<std::mutex> lk(mu);
    cv_done.wait(lk, [&] { return remaining == 0; });
  }
  ~Pool() {
    { std::lock_guard<std::mutex> lk(mu); quit = true; }
    cv.notify_all();
    for (auto& t : th) t.join();
  }
};
Pool& pool() { static Pool* p = new Pool(); return *p; }   // leaked on purpose: no join at exit
std::mutex launch_mu;
}  // namespace

void emu_run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  std::lock_guard<std::mutex> g(launch_mu);
  const int n = (int)(block.x * block.y * block.z);
  blockDim = EmuIdx{block.x, block.y, block.z};
  gridDim = EmuIdx{grid.x, grid.y, grid.z};
  std::vector<unsigned char> sm(smem + 64);
  emu_dyn_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sm.data()) + 63) & ~uintptr_t(63));
  const int nw = (n + 31) / 32;
  std::vector<EmuWarp> warps(nw);
  std::vector<std::unique_ptr<std::barrier<>>> wb;
  for (int w = 0; w < nw; ++w) {
    wb.emplace_back(new std::barrier<>(std::min(32, n - 32 * w)));
    warps[w].bar = wb.back().get();
  }
  emu_warps = warps.data();
  std::barrier<> bar(n);
  emu_block_barrier = &bar;
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) {
        // dynamic shared memory is uninitialised on the device; poison it here so reads of
        // unwritten cells show up as NaNs instead of lucky zeros.
        memset(emu_dyn_smem, 0xff, smem);
        pool().run_block(n, EmuIdx{x, y, z}, body);
      }
  emu_block_barrier = nullptr;
  emu_warps = nullptr;
  emu_dyn_smem = nullptr;
}

// ---------------------------------------------------------------- mbarrier stand-in (async_copy.cuh)
// One record per barrier address: pending arrivals, expected transaction bytes and the phase
// bit. A phase completes when both reach zero, as on the device.
#include <unordered_map>
namespace acq {
namespace {
struct MbarState { int init = 0, pending = 0; long long tx = 0; unsigned phase = 0; };
std::mutex mbar_mu;
std::unordered_map<const void*, MbarState> mbar_tab;
void mbar_settle(MbarState& s) {
  if (s.pending == 0 && s.tx == 0) { s.phase ^= 1u; s.pending = s.init; }
}
}  // namespace
void emu_mbar_init(unsigned long long* bar, int count) {
  std::lock_guard<std::mutex> g(mbar_mu);
  MbarState s; s.init = s.pending = count;
  mbar_tab[bar] = s;
}
void emu_mbar_arrive(unsigned long long* bar, long long tx) {
  std::lock_guard<std::mutex> g(mbar_mu);
  MbarState& s = mbar_tab.at(bar);
  s.tx += tx;
  s.pending -= 1;
  mbar_settle(s);
}
void emu_mbar_complete_tx(unsigned long long* bar, long long bytes) {
  std::lock_guard<std::mutex> g(mbar_mu);
  MbarState& s = mbar_tab.at(bar);
  s.tx -= bytes;
  mbar_settle(s);
}
// bar.sync id, count: `count` threads of the running block meet at barrier `id`
void emu_named_barrier(int id, int count) {
  static std::mutex mu;
  static int arrived[16] = {0};
  static long gen[16] = {0};
  long my;
  {
    std::lock_guard<std::mutex> g(mu);
    my = gen[id & 15];
    if (++arrived[id & 15] == count) { arrived[id & 15] = 0; ++gen[id & 15]; return; }
  }
  for (;;) {
    { std::lock_guard<std::mutex> g(mu); if (gen[id & 15] != my) return; }
    std::this_thread::yield();
  }
}
bool emu_mbar_test(unsigned long long* bar, unsigned parity) {
  std::lock_guard<std::mutex> g(mbar_mu);
  return mbar_tab.at(bar).phase != (parity & 1u);
}
}  // namespace acq
