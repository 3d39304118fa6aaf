// k_stats.cu -- the byte-level part of Stats.Call with --all:
//   Q20/Q30 base counts and gap counts     bigseqkit-lib/stats.go:90-102
// (the length histogram itself comes from the record index: lengths -> sort -> run-length encode)
#include "kernels.h"

namespace bsk {
namespace k {

// persistent warps, one record per warp iteration, 16-byte vector loads where aligned
__global__ void __launch_bounds__(256) k_stats_qual_gap(RecViews v, const u8 *__restrict__ gap, int fq_offset, int fastq,
                                                        DevStatus *st) {
  const u32 lane = threadIdx.x & 31;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  unsigned long long q20 = 0, q30 = 0, gaps = 0;
  const int t20 = fq_offset + 20, t30 = fq_offset + 30;
  for (u32 r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < v.n_rec; r += nwarps) {
    if (fastq) {
      const u8 *q = v.qualb + v.qual_off[r];
      const u32 ql = v.qual_len[r];
      for (u32 i = lane; i < ql; i += 32) {
        const int c = (int)q[i];
        if (c >= t20) {
          q20++;
          if (c >= t30) q30++;
        }
      }
    }
    const u8 *s = v.seqb + v.seq_off[r];
    const u32 sl = v.seq_len[r];
    for (u32 i = lane; i < sl; i += 32) gaps += gap[s[i]];
  }
  for (int off = 16; off > 0; off >>= 1) {
    q20 += __shfl_xor_sync(0xffffffffu, q20, off);
    q30 += __shfl_xor_sync(0xffffffffu, q30, off);
    gaps += __shfl_xor_sync(0xffffffffu, gaps, off);
  }
  if (lane == 0) {
    if (q20) atomicAdd((unsigned long long *)&st->counters[0], q20);
    if (q30) atomicAdd((unsigned long long *)&st->counters[1], q30);
    if (gaps) atomicAdd((unsigned long long *)&st->counters[2], gaps);
  }
}

void stats_qual_gap(RecViews v, const u8 *gap, int fq_offset, int fastq, DevStatus *st, cudaStream_t s) {
  if (!v.n_rec) return;
  u32 blocks = (v.n_rec + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  BSK_LAUNCH(k_stats_qual_gap, blocks, 256, 0, s, v, gap, fq_offset, fastq, st);
}

}  // namespace k
}  // namespace bsk
