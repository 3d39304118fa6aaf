// cuda_emu.cpp -- DEVELOPMENT TOOL ONLY (see cuda_emu.h).
#include "cuda_emu.h"
namespace emu {
Block *g_block = nullptr;
thread_local dim3 t_threadIdx, t_blockIdx;
dim3 g_blockDim, g_gridDim;

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
  std::vector<std::thread> th;
  for (unsigned t = 0; t < T; t++)
    th.emplace_back([&, t]() {
      for (unsigned b = 0; b < grid.x; b++) {
        t_blockIdx = dim3(b);
        t_threadIdx = dim3(t);
        body();
        blk.bar->arrive_and_wait();
      }
    });
  for (auto &x : th) x.join();
  g_block = nullptr;
}
}  // namespace emu
