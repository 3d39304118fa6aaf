// prims.cu -- thin wrappers over cub::Device* (plumbing: scans over per-record
// lengths, sorts of (key, index) pairs).  -DBSK_EMU builds use std:: algorithms.
#include "prims.h"

#ifndef BSK_EMU
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>
#else
#include <algorithm>
#include <numeric>
#include <vector>
#endif

namespace bsk {
namespace prim {

#ifndef BSK_EMU
#define BSK_CUB(call_with_tmp)                              \
  do {                                                      \
    void *d_tmp = nullptr;                                  \
    size_t bytes = 0;                                       \
    BSK_CUDA(call_with_tmp);                                \
    tmp.reserve(bytes ? bytes : 1);                         \
    d_tmp = tmp.p;                                          \
    BSK_CUDA(call_with_tmp);                                \
  } while (0)

void excl_scan_u64(const uint64_t *in, uint64_t *out, size_t n, DevBuf &tmp, cudaStream_t s) {
  if (!n) return;
  BSK_CUB(cub::DeviceScan::ExclusiveSum(d_tmp, bytes, in, out, n, s));
}
struct ToU64 {
  __host__ __device__ uint64_t operator()(uint32_t x) const { return (uint64_t)x; }
};
void excl_scan_u32_to_u64(const uint32_t *in, uint64_t *out, size_t n, DevBuf &tmp, cudaStream_t s) {
  if (!n) return;
  thrust::transform_iterator<ToU64, const uint32_t *> it(in, ToU64());
  BSK_CUB(cub::DeviceScan::ExclusiveSum(d_tmp, bytes, it, out, n, s));
}
void excl_scan_u32(const uint32_t *in, uint32_t *out, size_t n, DevBuf &tmp, cudaStream_t s) {
  if (!n) return;
  BSK_CUB(cub::DeviceScan::ExclusiveSum(d_tmp, bytes, in, out, n, s));
}
void select_flagged_u64(const uint64_t *in, const uint8_t *flags, uint64_t *out, uint32_t *d_count, size_t n, DevBuf &tmp,
                        cudaStream_t s) {
  if (!n) { BSK_CUDA(cudaMemsetAsync(d_count, 0, 4, s)); return; }
  BSK_CUB(cub::DeviceSelect::Flagged(d_tmp, bytes, in, flags, out, d_count, n, s));
}
void sort_u32(const uint32_t *in, uint32_t *out, size_t n, DevBuf &tmp, cudaStream_t s) {
  if (!n) return;
  BSK_CUB(cub::DeviceRadixSort::SortKeys(d_tmp, bytes, in, out, n, 0, 32, s));
}
void rle_u32(const uint32_t *in, uint32_t *uniq, uint32_t *counts, uint32_t *d_runs, size_t n, DevBuf &tmp, cudaStream_t s) {
  if (!n) { BSK_CUDA(cudaMemsetAsync(d_runs, 0, 4, s)); return; }
  BSK_CUB(cub::DeviceRunLengthEncode::Encode(d_tmp, bytes, in, uniq, counts, d_runs, n, s));
}
void sort_pairs_u64_u64(const uint64_t *kin, uint64_t *kout, const uint64_t *vin, uint64_t *vout, size_t n, int begin_bit,
                        int end_bit, DevBuf &tmp, cudaStream_t s) {
  if (!n) return;
  BSK_CUB(cub::DeviceRadixSort::SortPairs(d_tmp, bytes, kin, kout, vin, vout, n, begin_bit, end_bit, s));
}
#else
void excl_scan_u64(const uint64_t *in, uint64_t *out, size_t n, DevBuf &, cudaStream_t) {
  uint64_t a = 0;
  for (size_t i = 0; i < n; i++) { uint64_t v = in[i]; out[i] = a; a += v; }
}
void excl_scan_u32_to_u64(const uint32_t *in, uint64_t *out, size_t n, DevBuf &, cudaStream_t) {
  uint64_t a = 0;
  for (size_t i = 0; i < n; i++) { uint64_t v = in[i]; out[i] = a; a += v; }
}
void excl_scan_u32(const uint32_t *in, uint32_t *out, size_t n, DevBuf &, cudaStream_t) {
  uint32_t a = 0;
  for (size_t i = 0; i < n; i++) { uint32_t v = in[i]; out[i] = a; a += v; }
}
void select_flagged_u64(const uint64_t *in, const uint8_t *flags, uint64_t *out, uint32_t *d_count, size_t n, DevBuf &,
                        cudaStream_t) {
  uint32_t c = 0;
  for (size_t i = 0; i < n; i++) if (flags[i]) out[c++] = in[i];
  *d_count = c;
}
void sort_u32(const uint32_t *in, uint32_t *out, size_t n, DevBuf &, cudaStream_t) {
  std::vector<uint32_t> v(in, in + n);
  std::sort(v.begin(), v.end());
  std::copy(v.begin(), v.end(), out);
}
void rle_u32(const uint32_t *in, uint32_t *uniq, uint32_t *counts, uint32_t *d_runs, size_t n, DevBuf &, cudaStream_t) {
  uint32_t r = 0;
  for (size_t i = 0; i < n;) {
    size_t j = i;
    while (j < n && in[j] == in[i]) j++;
    uniq[r] = in[i]; counts[r] = (uint32_t)(j - i); r++;
    i = j;
  }
  *d_runs = r;
}
template <class V>
static void sort_pairs_impl(const uint64_t *kin, uint64_t *kout, const V *vin, V *vout, size_t n, int begin_bit, int end_bit) {
  std::vector<size_t> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  uint64_t mask = end_bit >= 64 ? ~0ull : ((1ull << end_bit) - 1);
  mask &= ~((1ull << begin_bit) - 1);
  std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return (kin[a] & mask) < (kin[b] & mask); });
  std::vector<uint64_t> k2(n);
  std::vector<V> v2(n);
  for (size_t i = 0; i < n; i++) { k2[i] = kin[idx[i]]; v2[i] = vin[idx[i]]; }
  std::copy(k2.begin(), k2.end(), kout);
  std::copy(v2.begin(), v2.end(), vout);
}
void sort_pairs_u64_u64(const uint64_t *kin, uint64_t *kout, const uint64_t *vin, uint64_t *vout, size_t n, int begin_bit,
                        int end_bit, DevBuf &, cudaStream_t) { sort_pairs_impl(kin, kout, vin, vout, n, begin_bit, end_bit); }
#endif

__global__ void k_copy_small(unsigned char *dst, const unsigned char *src, uint32_t n) {
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
}
void copy_small(void *dst, const void *src, uint32_t nbytes, cudaStream_t s) {
  if (!nbytes) return;
  BSK_LAUNCH_FLAT(k_copy_small, 1, nbytes < 128 ? 32 : 128, 0, s, static_cast<unsigned char *>(dst),
                  static_cast<const unsigned char *>(src), nbytes);
}
}  // namespace prim
}  // namespace bsk
