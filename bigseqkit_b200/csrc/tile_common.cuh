// tile_common.cuh -- front end shared by the TMA-staged tile kernels for short records (k_stats_tile.cu; the same
// code is inlined in k_fastq_inplace.cu, which came first): tile geometry, ring refill, edge fill and the newline
// scan that turns a staged region into a list of line starts with record-start flags.
//
//   PlainFile split + ReadFixer   bigseqkit/helper.go:148-178, bigseqkit-lib/helper.go:41-66
//   record-start rule: FASTA  a line that starts with '>'
//                                   FASTQ  a line that starts with '@' unless the line before it is a bare "+"
#pragma once
#include "kernels.h"
#include "tma.cuh"

namespace bsk {
namespace k {
namespace tile {

constexpr u32 H = 4096;   // halo bytes (longest record the tile kernels accept, roughly)
constexpr u32 PRE = 16;   // look-behind bytes in front of the tile

// CTA shape: NT threads, every lane scans CPL consecutive 16-byte chunks, so tile + halo = NT * CPL * 16 bytes
template <u32 NT_, u32 CPL_, u32 CTAS_, u32 NSTAGE_, u32 LCAP_>
struct Geo {
  static constexpr u32 NT = NT_, CPL = CPL_, CTAS = CTAS_, NSTAGE = NSTAGE_, LCAP = LCAP_;
  static constexpr u32 NWARP = NT / 32;
  static constexpr u32 T = NT * CPL * 16 - H;  // tile bytes
  static constexpr u32 STAGE = PRE + T + H + 16;
  static constexpr u32 LITER = (LCAP + NT) / NT;  // passes of the CTA over the line list
  static_assert(T % 16 == 0 && STAGE % 16 == 0 && T + H < 32768 && NWARP <= 16 && (CPL == 3 || CPL == 4), "tile geometry");
};

// flags (0x80 per byte) of the bytes of w that equal '\n'; exact for every byte value.  (x | 0x80) - 1 keeps bit 7 of a
// byte unless its low seven bits are zero, and never borrows across bytes; bit 7 of w ^ '\n' is bit 7 of w.
__device__ __forceinline__ u32 nl_flags(u32 w) {
  const u32 t = ((w ^ 0x0a0a0a0au) | 0x80808080u) - 0x01010101u;
  return ~(t | w) & 0x80808080u;
}

// global byte range [g0, g1) that one bulk load brings for `tile` (whole 16-byte chunks of the file only)
template <class G>
__device__ __forceinline__ bool bulk_range(u32 tile, u32 n16, u32 &g0, u32 &g1) {
  const u32 t0 = tile * G::T;  // n < 4 GiB - 1 MiB (engine.h kMaxBlockBytes): t0 + T + H does not wrap
  g0 = tile ? t0 - PRE : 0u;
  g1 = t0 + G::T + H < n16 ? t0 + G::T + H : n16;
  return g1 > g0;
}
template <class G>
__device__ __forceinline__ void issue_load(const u8 *in, u32 n16, u32 tile, u8 *stage, u64 *bar) {
  u32 g0, g1;
  if (bulk_range<G>(tile, n16, g0, g1)) {
    tma::mbar_expect_tx(bar, g1 - g0);
    tma::bulk_load(stage + (g0 + PRE - tile * G::T), in + g0, g1 - g0, bar);
  }
}
// bytes the bulk copy did not bring: the look-behind of tile 0, the ragged tail of the file, '\n' padding.
// Contains a CTA barrier when it does anything (uniform).
template <class G>
__device__ __forceinline__ void fill_edges(const u8 *in, u32 n, u32 tile, u8 *stage) {
  const u32 n16 = n & ~15u, t0 = tile * G::T;
  if (tile == 0 || t0 + G::T + H > n16) {
    for (u32 i = threadIdx.x; i < G::STAGE; i += G::NT) {
      const bool before = t0 + i < PRE;  // only tile 0
      const u32 g = t0 + i - PRE;
      if (before || g >= n16) stage[i] = (!before && g < n) ? in[g] : (u8)'\n';
    }
    __syncthreads();
  }
}

// Newline scan of region bytes [0, slim) -> ls[0 .. n_lines]: start of every line (ls[0] = 0), bit 15 set when the
// line opens a record; ls[n_lines] is the start of the line after the last terminated one (or lim + 1 behind an
// unterminated last line at the end of the file).  wtot: NWARP words of scratch.  `pre` runs on thread 0 before
// the first of the two CTA barriers.  Returns n_lines, or 0xffffffff when the list would overflow (uniform).
template <class G, class Pre>
__device__ __forceinline__ u32 scan_lines(const u8 *d, u32 lim, u32 slim, bool eof, u32 t0, u32 tile, bool fq, u16 *ls,
                                          u32 *wtot, Pre pre) {
  constexpr u32 CPL = G::CPL, NWARP = G::NWARP;
  const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const u8 marker = fq ? '@' : '>';
  const u32 span = (warp * 32u + lane) * (CPL * 16u);
  u32 mlo = 0, mhi = 0;
  if (warp * (32u * CPL * 16u) < slim) {
    // the 0x80 flag bytes of a chunk are packed into a position-ordered 16-bit mask with four IDP.4A
    u32 m16[CPL];
#pragma unroll
    for (u32 j = 0; j < CPL; j++) {
      const uint4 v = *reinterpret_cast<const uint4 *>(d + span + j * 16u);
      u32 lo = __dp4a(nl_flags(v.x), 0x08040201u, 0u);
      lo = __dp4a(nl_flags(v.y), 0x80402010u, lo);
      u32 hi = __dp4a(nl_flags(v.z), 0x08040201u, 0u);
      hi = __dp4a(nl_flags(v.w), 0x80402010u, hi);
      m16[j] = (lo >> 7) | (hi << 1);
    }
    mlo = m16[0] | (m16[1] << 16);
    mhi = m16[2];
    if (CPL == 4) mhi |= m16[CPL - 1] << 16;
    if (span + CPL * 16u > slim) {  // bytes past the scanned range (or the end of the file) do not count
      const u32 valid = slim > span ? slim - span : 0u;
      if (valid < 32u) { mlo &= (1u << valid) - 1u; mhi = 0; }
      else if (valid < CPL * 16u) mhi &= (1u << (valid - 32u)) - 1u;
    }
  }
  const u32 cnt = __popc(mlo) + __popc(mhi);
  u32 inc = cnt;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const u32 y = __shfl_up_sync(0xffffffffu, inc, off);
    if ((int)lane >= off) inc += y;
  }
  if (lane == 31) wtot[warp] = inc;
  if (tid == 0) pre();
  __syncthreads();
  u32 base, n_nl;
  {
    u32 x = (lane & 15u) < NWARP ? wtot[lane & 15u] : 0u;
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
      const u32 y = __shfl_up_sync(0xffffffffu, x, off, 16);
      if ((int)(lane & 15u) >= off) x += y;
    }
    n_nl = __shfl_sync(0xffffffffu, x, 15);
    base = __shfl_sync(0xffffffffu, x, (warp + 15u) & 15u);
    if (warp == 0) base = 0;
  }
  const bool virt = eof && slim == lim && lim > 0 && d[lim - 1] != '\n';  // unterminated last line
  if (n_nl + 2 > G::LCAP) return 0xffffffffu;
  u32 k = base + inc - cnt + 1;  // ls[k] = start of the line after the (k-1)-th newline
  while (mlo | mhi) {
    u32 t;
    if (mlo) { t = (u32)__ffs((int)mlo) - 1u; mlo &= mlo - 1u; }
    else { t = 32u + (u32)__ffs((int)mhi) - 1u; mhi &= mhi - 1u; }
    const u32 p = span + t + 1u;
    u32 rs = 0;
    if (p < lim && d[p] == marker) rs = (fq && t0 + p >= 3u && d[(int)p - 3] == '\n' && d[(int)p - 2] == '+') ? 0u : 0x8000u;
    ls[k++] = (u16)(p | rs);
  }
  if (tid == 0) {
    u32 rs0 = 0;
    if (lim > 0 && d[0] == marker) {
      if (tile == 0) rs0 = 0x8000u;
      else if (d[-1] == '\n' && !(fq && d[-3] == '\n' && d[-2] == '+')) rs0 = 0x8000u;
    }
    ls[0] = (u16)rs0;
    if (virt) ls[n_nl + 1] = (u16)(lim + 1);
  }
  __syncthreads();
  return n_nl + (virt ? 1u : 0u);
}

}  // namespace tile
}  // namespace k
}  // namespace bsk
