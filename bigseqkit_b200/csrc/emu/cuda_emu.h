// cuda_emu.h -- DEVELOPMENT TOOL ONLY.  A tiny host emulation of the CUDA
// execution model (blocks run one after another; the threads of a block are OS
// threads with real barriers) so that the kernel LOGIC of csrc/*.cu can be run
// under gdb / ASan in the GPU-less build container before GPU minutes are
// spent.  It is compiled only into tools/emu builds (-DBSK_EMU); the product
// library libbsk.so is always built by nvcc for sm_100a and never contains or
// falls back to this code.
#pragma once
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct ulonglong2 { unsigned long long x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __constant__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

namespace emu {
struct Block {
  unsigned nthreads = 0;
  std::unique_ptr<std::barrier<>> bar;
  std::vector<std::unique_ptr<std::barrier<>>> wbar;
  std::vector<uint64_t> wslot;  // 32 slots per warp
  std::vector<uint8_t> dyn;     // dynamic shared memory
  std::atomic<int> red{0};      // __syncthreads_or scratch
};
extern Block *g_block;
extern thread_local dim3 t_threadIdx, t_blockIdx;
extern dim3 g_blockDim, g_gridDim;
void launch(dim3 grid, dim3 block, size_t smem, bool coop, const std::function<void()> &body);
inline void sync_block() { if (g_block->bar) g_block->bar->arrive_and_wait(); }
inline void sync_warp() { if (!g_block->wbar.empty()) g_block->wbar[t_threadIdx.x >> 5]->arrive_and_wait(); }
inline uint64_t *wslots() { return &g_block->wslot[(t_threadIdx.x >> 5) * 32]; }
}  // namespace emu

#define threadIdx emu::t_threadIdx
#define blockIdx emu::t_blockIdx
#define blockDim emu::g_blockDim
#define gridDim emu::g_gridDim
#define warpSize 32

static inline void __syncthreads() { emu::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::sync_warp(); }
static inline int __syncthreads_or(int p) {
  if (p) emu::g_block->red.fetch_or(1);
  emu::sync_block();
  const int r = emu::g_block->red.load();
  emu::sync_block();
  if (emu::t_threadIdx.x == 0) emu::g_block->red.store(0);
  emu::sync_block();
  return r;
}
static inline int __syncthreads_count(int p) {
  static std::atomic<int> cnt{0};  // one block runs at a time in the emulator
  if (p) cnt.fetch_add(1);
  emu::sync_block();
  const int r = cnt.load();
  emu::sync_block();
  if (emu::t_threadIdx.x == 0) cnt.store(0);
  emu::sync_block();
  return r;
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }

template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  static_assert(sizeof(T) <= 8, "shfl");
  uint64_t *s = emu::wslots();
  int lane = threadIdx.x & 31;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  s[lane] = raw;
  emu::sync_warp();
  int base = lane & ~(width - 1);
  uint64_t r = s[base + (src & (width - 1))];
  emu::sync_warp();
  T out;
  memcpy(&out, &r, sizeof(T));
  return out;
}
template <class T>
static inline T __shfl_up_sync(unsigned m, T v, unsigned d, int width = 32) {
  int lane = threadIdx.x & 31;
  T r = __shfl_sync(m, v, lane >= (int)d ? lane - (int)d : lane, width);
  return lane >= (int)d ? r : v;
}
template <class T>
static inline T __shfl_down_sync(unsigned m, T v, unsigned d, int width = 32) {
  int lane = threadIdx.x & 31;
  T r = __shfl_sync(m, v, lane + (int)d < 32 ? lane + (int)d : lane, width);
  return lane + (int)d < 32 ? r : v;
}
template <class T>
static inline T __shfl_xor_sync(unsigned m, T v, int x, int width = 32) {
  int lane = threadIdx.x & 31;
  return __shfl_sync(m, v, lane ^ x, width);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  uint64_t *s = emu::wslots();
  int lane = threadIdx.x & 31;
  s[lane] = pred ? 1 : 0;
  emu::sync_warp();
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (unsigned)(s[i] & 1) << i;
  emu::sync_warp();
  return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, p) == 0xffffffffu; }
static inline unsigned __activemask() { return 0xffffffffu; }

static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __brev(unsigned x) {
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
  return r;
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
  uint64_t v = ((uint64_t)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) {
    unsigned sel = (s >> (4 * i)) & 0xf;
    unsigned byte = (unsigned)(v >> (8 * (sel & 7))) & 0xff;
    if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
    r |= byte << (8 * i);
  }
  return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (unsigned)(v >> (sh & 31));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (unsigned)((v << (sh & 31)) >> 32);
}
static inline unsigned __vcmpeq4(unsigned a, unsigned b) {
  unsigned r = 0;
  for (int i = 0; i < 4; i++)
    if (((a >> (8 * i)) & 0xff) == ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i);
  return r;
}
static inline unsigned __vcmpgeu4(unsigned a, unsigned b) {
  unsigned r = 0;
  for (int i = 0; i < 4; i++)
    if (((a >> (8 * i)) & 0xff) >= ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i);
  return r;
}
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
  for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
  return c;
}
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
  return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }

// atomics (all seq_cst; blocks run sequentially, threads of a block concurrently)
template <class T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicAnd(T *p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicCAS(T *p, T cmp, T v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}
template <class T> static inline T atomicMin(T *p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <class T> static inline T atomicMax(T *p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}

// ---- runtime API subset -------------------------------------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef struct emuEvent { double t; } *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaHostAllocPortable = 1, cudaHostAllocMapped = 2, cudaEventDefault = 0, cudaEventDisableTiming = 2 };
static inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { if (n) memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { if (n) memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emuEvent{0}; return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new emuEvent{0}; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
struct cudaPointerAttributes { int type; };
enum { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) { a->type = cudaMemoryTypeHost; return 0; }
struct cudaDeviceProp { int multiProcessorCount; };
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { p->multiProcessorCount = 4; return 0; }

#define BSK_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch(dim3(grid), dim3(block), (smem), true, [=]() { kernel(__VA_ARGS__); })
#define BSK_LAUNCH_FLAT(kernel, grid, block, smem, stream, ...) \
  emu::launch(dim3(grid), dim3(block), (smem), false, [=]() { kernel(__VA_ARGS__); })
#define BSK_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(emu::g_block->dyn.data())
