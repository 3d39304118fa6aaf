// prims.h -- device-wide scan / sort / select plumbing (CUB in the product build).
#pragma once
#include <cstdint>

#include "devbuf.h"

namespace bsk {
namespace prim {
void excl_scan_u64(const uint64_t *in, uint64_t *out, size_t n, DevBuf &tmp, cudaStream_t s);
void excl_scan_u32_to_u64(const uint32_t *in, uint64_t *out, size_t n, DevBuf &tmp, cudaStream_t s);
void excl_scan_u32(const uint32_t *in, uint32_t *out, size_t n, DevBuf &tmp, cudaStream_t s);
void select_flagged_u64(const uint64_t *in, const uint8_t *flags, uint64_t *out, uint32_t *d_count, size_t n, DevBuf &tmp,
                        cudaStream_t s);
void sort_u32(const uint32_t *in, uint32_t *out, size_t n, DevBuf &tmp, cudaStream_t s);
void rle_u32(const uint32_t *in, uint32_t *uniq, uint32_t *counts, uint32_t *d_runs, size_t n, DevBuf &tmp, cudaStream_t s);
// stable LSD radix sort on key bits [begin_bit, end_bit)
void sort_pairs_u64_u64(const uint64_t *kin, uint64_t *kout, const uint64_t *vin, uint64_t *vout, size_t n, int begin_bit,
                        int end_bit, DevBuf &tmp, cudaStream_t s);
// A few bytes device -> page-locked host memory (or device -> device) by a kernel instead of a copy engine: a small
// cudaMemcpyAsync queues behind the block-sized copies of the other streams on the same engine (FIFO), which ties the
// compute stream of block i to the transfers of blocks i-1 / i+1.  dst may be any cudaHostAlloc'd address (UVA).
void copy_small(void *dst, const void *src, uint32_t nbytes, cudaStream_t s);
}  // namespace prim
}  // namespace bsk
