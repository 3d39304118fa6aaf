// k_seq.cu -- record-level pieces of `seq` that are not plain formatting:
//   Seq.RemoveGapsInplace        call site bigseqkit-lib/seq.go:129-131
//   length / average-quality filters       bigseqkit-lib/seq.go:133-149 (Seq.AvgQual of bio v0.7.0)
// Reverse / complement / dna2rna / case changes are byte maps and are folded into
// the emitter (k_emit.cu) as a mirrored index plus one 256-entry table.
#include <cmath>

#include "kernels.h"

namespace bsk {
namespace k {

// one warp per record; ballot/popc stream compaction of the bytes that are not gap letters
__global__ void __launch_bounds__(256) k_remove_gaps(RecViews v, const u8 *__restrict__ gap, u8 *__restrict__ seq_out,
                                                     u8 *__restrict__ qual_out, u32 *__restrict__ new_len, int has_qual) {
  const u32 lane = threadIdx.x & 31;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  for (u32 r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < v.n_rec; r += nwarps) {
    const u32 len = v.seq_len[r], so = v.seq_off[r];
    const bool q = has_qual && v.qual_len[r] == len;
    const u32 qo = q ? v.qual_off[r] : 0;
    u32 w = 0;
    for (u32 base = 0; base < len; base += 32) {
      const u32 i = base + lane;
      u8 c = 0;
      bool keepb = false;
      if (i < len) {
        c = v.seqb[so + i];
        keepb = !gap[c];
      }
      const u32 bal = __ballot_sync(0xffffffffu, keepb);
      if (keepb) {
        const u32 rank = (u32)__popc(bal & ((1u << lane) - 1u));
        seq_out[so + w + rank] = c;
        if (q) qual_out[qo + w + rank] = v.qualb[qo + i];
      }
      w += (u32)__popc(bal);
    }
    if (lane == 0) new_len[r] = w;
  }
}

// one thread per record.  AvgQual: sequential double sum of 10^(-(q-base)/10) in input
// order (same order as the scalar reference, so the sum is bit-identical), table made on the host.
__global__ void k_seq_filter(RecViews v, int min_len, int max_len, double min_qual, double max_qual,
                             const double *__restrict__ qual_pow, u8 *__restrict__ keep) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= v.n_rec) return;
  const u32 len = v.seq_len[r];
  bool k = true;
  if (min_len > 0 && (long long)len < (long long)min_len) k = false;
  if (max_len > 0 && (long long)len > (long long)max_len) k = false;
  if (k && (min_qual > 0 || max_qual > 0)) {
    double avg = 0;
    const u32 ql = v.qual_len[r];
    if (ql > 0) {
      const u8 *q = v.qualb + v.qual_off[r];
      double sum = 0;
      for (u32 i = 0; i < ql; i++) sum += qual_pow[q[i]];
      avg = -10.0 * log10(sum / (double)ql);
    }
    if (min_qual > 0 && avg < min_qual) k = false;
    if (max_qual > 0 && avg >= max_qual) k = false;
  }
  keep[r] = k ? 1 : 0;
}

void remove_gaps(RecViews v, const u8 *gap, u8 *seq_out, u8 *qual_out, u32 *new_len, int has_qual, cudaStream_t s) {
  if (!v.n_rec) return;
  u32 blocks = (v.n_rec + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  BSK_LAUNCH(k_remove_gaps, blocks, 256, 0, s, v, gap, seq_out, qual_out, new_len, has_qual);
}
void seq_filter(RecViews v, int min_len, int max_len, double min_qual, double max_qual, const double *qual_pow, u8 *keep,
                cudaStream_t s) {
  if (v.n_rec) BSK_LAUNCH_FLAT(k_seq_filter, (v.n_rec + 255) / 256, 256, 0, s, v, min_len, max_len, min_qual, max_qual, qual_pow, keep);
}

}  // namespace k
}  // namespace bsk
