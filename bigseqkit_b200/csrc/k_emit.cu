// k_emit.cu -- generic record formatter, output-driven.
//
// Replaces the per-record buffer assembly of the reference:
//   SeqTransform.Call writer part        bigseqkit-lib/seq.go:151-265
//   fastx.Record.Format(width)           (bio v0.7.0; call sites rmdup.go:86,214, grep.go:529, subseq.go:316)
//   wrapByteSlice                        bigseqkit-lib/helper.go:81-117
//   FileStore "element + \n"             bigseqkit-lib/helper.go:441-451
//
// Every thread owns 16 consecutive OUTPUT bytes (one aligned 16-byte store) and
// gathers their source bytes: short reads and megabase contigs take the same path
// and stores are always fully coalesced.
#include "kernels.h"

namespace bsk {
namespace k {

__device__ __forceinline__ u32 wrap_len(u32 l, u32 w) { return (w < 1 || l == 0) ? l : l + (l - 1) / w; }

struct RecOut {
  u32 name_off, name_len, seq_off, seq_len, qual_off, qual_len;
  u32 np, ns, nq, wl;
};

__device__ __forceinline__ RecOut load_rec(const RecViews &v, const EmitCfg &c, u32 r) {
  RecOut o;
  o.name_off = v.name_off[r];
  o.name_len = v.name_len[r];
  o.seq_off = v.seq_off[r];
  o.seq_len = v.seq_len[r];
  o.qual_off = c.print_qual ? v.qual_off[r] : 0;
  o.qual_len = c.print_qual ? v.qual_len[r] : 0;
  o.np = c.print_name ? (c.marker ? 1u : 0u) + o.name_len + 1u : 0u;
  o.wl = wrap_len(o.seq_len, c.width);
  o.ns = c.print_seq ? o.wl + 1u : 0u;
  o.nq = c.print_qual ? (c.plus_line ? 2u : 0u) + o.qual_len + 1u : 0u;
  return o;
}

__global__ void k_out_len(RecViews v, EmitCfg c, const u8 *__restrict__ keep, u32 *__restrict__ out_len) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > v.n_rec) return;
  if (r == v.n_rec) { out_len[r] = 0; return; }
  if (keep && !keep[r]) { out_len[r] = 0; return; }
  RecOut o = load_rec(v, c, r);
  out_len[r] = o.np + o.ns + o.nq;
}

__device__ __forceinline__ u8 rec_byte(const RecViews &v, const EmitCfg &c, const RecOut &o, u32 p,
                                       const u8 *__restrict__ lut) {
  if (p < o.np) {
    if (c.marker) {
      if (p == 0) return c.marker;
      p--;
    }
    return p < o.name_len ? v.in[o.name_off + p] : (u8)'\n';
  }
  p -= o.np;
  if (p < o.ns) {
    if (p == o.wl) return '\n';
    u32 j = p;
    if (c.width > 0) {
      const u32 line = p / (c.width + 1u), col = p - line * (c.width + 1u);
      if (col == c.width) return '\n';
      j = line * c.width + col;
    }
    if (c.reverse) j = o.seq_len - 1u - j;
    const u8 b = v.seqb[o.seq_off + j];
    return lut ? lut[b] : b;
  }
  p -= o.ns;
  if (c.plus_line) {
    if (p == 0) return '+';
    if (p == 1) return '\n';
    p -= 2;
  }
  if (p < o.qual_len) return v.qualb[o.qual_off + (c.reverse ? o.qual_len - 1u - p : p)];
  return '\n';
}

__global__ void __launch_bounds__(256) k_emit(RecViews v, EmitCfg c, const u64 *__restrict__ off, u8 *__restrict__ out,
                                              u64 total, const u8 *__restrict__ lut) {
  const u64 o = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * 16ull;
  if (o >= total) return;
  u32 lo = 0, hi = v.n_rec;  // off[lo] <= o < off[hi]
  while (hi - lo > 1) {
    const u32 mid = lo + ((hi - lo) >> 1);
    if (off[mid] <= o) lo = mid;
    else hi = mid;
  }
  u32 r = lo;
  RecOut ro = load_rec(v, c, r);
  u64 rend = off[r + 1];
  u32 p = (u32)(o - off[r]);
  u32 w[4] = {0, 0, 0, 0};
#pragma unroll
  for (int b = 0; b < 16; b++) {
    const u64 pos = o + (u64)b;
    if (pos < total) {
      if (pos >= rend) {
        do {
          r++;
          rend = off[r + 1];
        } while (pos >= rend);
        ro = load_rec(v, c, r);
        p = 0;
      }
      w[b >> 2] |= (u32)rec_byte(v, c, ro, p, lut) << (8 * (b & 3));
      p++;
    }
  }
  *reinterpret_cast<uint4 *>(out + o) = make_uint4(w[0], w[1], w[2], w[3]);
}

void out_len(RecViews v, EmitCfg c, const u8 *keep, u32 *out_len_, cudaStream_t s) {
  BSK_LAUNCH_FLAT(k_out_len, (v.n_rec + 1 + 255) / 256, 256, 0, s, v, c, keep, out_len_);
}
void emit(RecViews v, EmitCfg c, const u64 *out_off, u8 *out, u64 total, const u8 *lut, cudaStream_t s) {
  if (!total) return;
  const u64 threads = (total + 15) / 16;
  BSK_LAUNCH_FLAT(k_emit, (u32)((threads + 255) / 256), 256, 0, s, v, c, out_off, out, total, lut);
}

}  // namespace k
}  // namespace bsk
