// tma.cuh -- 1-D bulk async copies (TMA, cp.async.bulk) + mbarrier helpers for sm_100a.
//
// The tile kernels stream FASTA/FASTQ bytes HBM -> shared memory -> HBM with the copy engine so
// that no thread spends issue slots on global loads/stores: one elected thread arms an mbarrier
// with the byte count and issues the bulk load; consumers wait on the barrier's phase parity;
// results leave through bulk stores grouped per issuing thread.
//
// -DBSK_EMU (development emulator, never shipped): the copies are plain memcpy's that complete
// at issue time, the mbarrier is a phase counter.
#pragma once
#include "kernels.h"

namespace bsk {
namespace tma {

#ifndef BSK_EMU
__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
// make the initialised barriers visible to the async proxy (call once after mbar_init, before __syncthreads)
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// order this thread's generic-proxy shared-memory writes before later async-proxy (bulk store) reads
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// global -> shared, 16-byte aligned addresses, bytes % 16 == 0; completion is signalled on `bar`
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, u32 bytes, u64 *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(u64 *bar, u32 parity) {
  u32 ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// shared -> global, 16-byte aligned addresses, bytes % 16 == 0; joins the calling thread's current bulk group
__device__ __forceinline__ void bulk_store(void *gdst, const void *smem_src, u32 bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_addr(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores of this thread have finished READING shared memory (the source may be reused)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all bulk stores of this thread are complete
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// split CTA barrier (named barrier `id`, `count` threads): producers arrive and go on, the consumer waits for all of them
__device__ __forceinline__ void named_arrive(u32 id, u32 count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_sync(u32 id, u32 count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

#else  // ---- emulator
static inline void mbar_init(u64 *bar, u32) { __atomic_store_n(bar, 0ull, __ATOMIC_SEQ_CST); }
static inline void fence_barrier_init() {}
static inline void fence_proxy_async() {}
static inline void mbar_expect_tx(u64 *, u32) {}
static inline void bulk_load(void *smem_dst, const void *gsrc, u32 bytes, u64 *bar) {
  memcpy(smem_dst, gsrc, bytes);
  __atomic_fetch_add(bar, 1ull, __ATOMIC_SEQ_CST);  // one load == one completed phase
}
static inline bool mbar_try_wait(u64 *bar, u32 parity) { return (__atomic_load_n(bar, __ATOMIC_SEQ_CST) & 1ull) != parity; }
static inline void mbar_wait(u64 *bar, u32 parity) {
  while (!mbar_try_wait(bar, parity)) std::this_thread::yield();
}
static inline void bulk_store(void *gdst, const void *smem_src, u32 bytes) { memcpy(gdst, smem_src, bytes); }
static inline void bulk_commit() {}
static inline void bulk_wait_read() {}
static inline void bulk_wait_all() {}
static inline void named_arrive(u32, u32) { __syncthreads(); }  // no overlap in the emulator: both sides meet at a full barrier
static inline void named_sync(u32, u32) { __syncthreads(); }
#endif

}  // namespace tma
}  // namespace bsk
