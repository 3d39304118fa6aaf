// cuda_emu.cpp -- DEVELOPMENT TOOL ONLY (see cuda_emu.h).
#include "cuda_emu.h"

#include <condition_variable>
#include <mutex>

namespace emu {
Block *g_block = nullptr;
thread_local dim3 t_threadIdx, t_blockIdx;
dim3 g_blockDim, g_gridDim;

namespace {
// Persistent worker pool: creating hundreds of OS threads for every kernel launch dominated the run time of the
// logic tests.  Workers sleep on a generation counter; a launch wakes the first T of them.
struct Pool {
  std::mutex mu;
  std::condition_variable cv_start, cv_done;
  std::vector<std::thread> workers;
  unsigned long generation = 0;
  unsigned active = 0, remaining = 0, grid = 0;
  const std::function<void()> *body = nullptr;
  Block *blk = nullptr;

  void worker(unsigned t) {
    unsigned long seen = 0;
    for (;;) {
      const std::function<void()> *fn;
      Block *b;
      unsigned g;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_start.wait(lk, [&] { return generation != seen && (t < active || generation == ~0ul); });
        if (generation == ~0ul) return;
        seen = generation;
        fn = body;
        b = blk;
        g = grid;
      }
      for (unsigned blkid = 0; blkid < g; blkid++) {
        t_blockIdx = dim3(blkid);
        t_threadIdx = dim3(t);
        (*fn)();
        b->bar->arrive_and_wait();
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        if (--remaining == 0) cv_done.notify_all();
      }
    }
  }

  void run(unsigned T, unsigned g, Block *b, const std::function<void()> &fn) {
    {
      std::lock_guard<std::mutex> lk(mu);
      while (workers.size() < T) {
        const unsigned t = (unsigned)workers.size();
        workers.emplace_back([this, t] { worker(t); });
      }
      active = T;
      remaining = T;
      grid = g;
      body = &fn;
      blk = b;
      generation++;
    }
    cv_start.notify_all();
    std::unique_lock<std::mutex> lk(mu);
    cv_done.wait(lk, [&] { return remaining == 0; });
  }

  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(mu);
      generation = ~0ul;
    }
    cv_start.notify_all();
    for (auto &w : workers) w.join();
  }
};
Pool &pool() {
  static Pool p;
  return p;
}
}  // namespace

void launch(dim3 grid, dim3 block, size_t smem, bool coop, const std::function<void()> &body) {
  Block blk;
  blk.nthreads = block.x;
  blk.dyn.assign(smem + 64, 0);
  g_block = &blk;
  g_blockDim = block;
  g_gridDim = grid;
  if (grid.x == 0) return;
  if (!coop || block.x == 1) {
    for (unsigned b = 0; b < grid.x; b++)
      for (unsigned t = 0; t < block.x; t++) {
        t_blockIdx = dim3(b);
        t_threadIdx = dim3(t);
        body();
      }
    g_block = nullptr;
    return;
  }
  unsigned T = block.x, nw = (T + 31) / 32;
  blk.bar.reset(new std::barrier<>(T));
  for (unsigned w = 0; w < nw; w++) {
    unsigned c = T - w * 32 < 32 ? T - w * 32 : 32;
    blk.wbar.emplace_back(new std::barrier<>(c));
  }
  blk.wslot.assign(nw * 32, 0);
  pool().run(T, grid.x, &blk, body);
  g_block = nullptr;
}
}  // namespace emu
