// k_fasta_tile.cu -- record table of a FASTA block in ONE streaming pass over the raw bytes, for the operators that
// read wrapped sequences in place (translate; same skeleton as k_locate_tile.cu, without the k-mer pass).
//
//   PlainFile split + ReadFixer   bigseqkit/helper.go:148-178, bigseqkit-lib/helper.go:41-66   records = lines that start with '>'
//   SeqParser.Read (FASTA)        bigseqkit-lib/helper.go:236-250   head = first line, sequence = the other lines joined
//
// Persistent CTAs walk 23 KiB tiles (+ 1 KiB look-behind) staged by 1-D TMA bulk loads.  Per tile:
//   * SWAR newline scan -> newline masks per lane, CTA prefix -> newlines per tile (a later prefix over the tiles turns a
//     raw byte position into a sequence coordinate: bases = bytes - newlines);
//   * every '>' that follows a newline is appended to the header list (position, newlines of its tile in front);
//   * line-shape check, local to every newline: with `width` > 0 every sequence line that is not the last of its record
//     must hold exactly `width` bytes and the last one at most that; with `width` == 0 every record has one sequence
//     line.  When the check holds for the whole block, base b of a record sits at seq_start + b + b / width, and the
//     consumers fetch bases with that arithmetic instead of a newline-squeezed copy;
//   * one flag byte per 96-byte span: bit j set when the sequence bytes of 16-byte chunk j are nothing but A, C, G, T
//     (header bytes and '\n' do not count; the consumers' fast path skips its own per-byte checks there).
// Reads N bytes, writes N / 96 + 8 bytes per tile + 16 bytes per record.
#include "kernels.h"
#include "tma.cuh"

namespace bsk {
namespace k {

namespace ft {
constexpr u32 NT = 256, SPAN = 96, REGION = NT * SPAN, LBL = 11, LB = LBL * SPAN, T = REGION - LB, NWARP = NT / 32;
constexpr u32 NSTAGE = 2;
constexpr u32 CTAS = 4;  // CTAs per SM (two 24 KiB stages each)
constexpr u32 AHEAD = 16;  // bytes staged behind the tile: the byte after the tile's last newline
struct Smem {
  u8 in[NSTAGE][REGION + AHEAD];
  u64 full[NSTAGE];
};
}  // namespace ft

__device__ __forceinline__ u32 ft_nl_flags(u32 w) {
  const u32 x = w ^ 0x0a0a0a0au;
  const u32 y = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
  return ~(y | x) & 0x80808080u;
}
// nonzero when some byte of w is neither an upper-case A / C / G / T nor '\n' (nl: 0x80 flags of the '\n' bytes of w)
__device__ __forceinline__ u32 ft_not_acgt(u32 w, u32 nl) {
  const u32 codes = (w >> 1) & 0x03030303u;              // A 0, C 1, T 2, G 3
  const u32 t = codes | (codes >> 4);
  const u32 sel = __byte_perm(t, 0u, 0x4420u);           // one selector nibble per byte
  const u32 expect = __byte_perm(0x47544341u, 0u, sel);  // "ACTG"[code]
  return (expect ^ w) & ~((nl >> 7) * 0xffu);
}

__global__ void __launch_bounds__(ft::NT, ft::CTAS) k_fasta_index_tile(FastaTileArgs a) {
  using namespace ft;
  BSK_DYN_SMEM(Smem, smp);
  Smem &sm = *smp;
  __shared__ u32 s_wtot[NWARP], s_wlast[NWARP];
  __shared__ u32 s_nl_lb, s_bad;
  const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const u32 n = a.n, n16 = n & ~15u;

  if (tid == 0) {
    for (u32 s = 0; s < NSTAGE; s++) tma::mbar_init(&sm.full[s], 1);
    tma::fence_barrier_init();
    s_bad = 0;
  }
  __syncthreads();
  auto bulk_range = [&](u32 tile, u32 &g0, u32 &g1) {
    const u32 t0 = tile * T;
    g0 = t0 >= LB ? t0 - LB : 0u;
    g1 = t0 + T + AHEAD < n16 ? t0 + T + AHEAD : n16;
    return g1 > g0;
  };
  auto issue = [&](u32 tile, u32 s) {
    u32 g0, g1;
    if (bulk_range(tile, g0, g1)) {
      tma::mbar_expect_tx(&sm.full[s], g1 - g0);
      tma::bulk_load(&sm.in[s][g0 + LB - tile * T], a.in + g0, g1 - g0, &sm.full[s]);
    }
  };
  if (tid == 0) {
    for (u32 p = 0; p < NSTAGE; p++) {
      const u32 tl = blockIdx.x + p * gridDim.x;
      if (tl < a.n_tiles) issue(tl, p);
    }
  }

  u32 it = 0;
  for (u32 tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, it++) {
    const u32 s = it % NSTAGE, parity = (it / NSTAGE) & 1u;
    const u32 t0 = tile * T;
    u8 *d = sm.in[s];                                  // region byte i == global byte t0 - LB + i
    const u32 lim = (n - t0 < T ? n - t0 : T) + LB;    // valid bytes of the region (look-behind included, look-ahead not)
    const u32 lim_a = (n - t0 < T + AHEAD ? n - t0 : T + AHEAD) + LB;  // ... with the look-ahead
    const bool eof = n - t0 <= T;                      // the file ends inside this tile
    {
      u32 g0, g1;
      if (bulk_range(tile, g0, g1)) tma::mbar_wait(&sm.full[s], parity);
    }
    if (t0 < LB || t0 + T + AHEAD > n16) {
      for (u32 i = tid; i < REGION + AHEAD; i += NT) {
        const bool before = t0 + i < LB;
        const u32 g = t0 + i - LB;
        if (before || g >= n16) d[i] = (!before && g < n) ? a.in[g] : (u8)'\n';
      }
      __syncthreads();
    }

    // ---- newline masks + clean-chunk bits of this lane's span
    const u32 span0 = tid * SPAN;
    u32 m[3], clean = 0;
    {
      u32 m16[6];
#pragma unroll
      for (u32 j = 0; j < 6; j++) {
        const uint4 v = *reinterpret_cast<const uint4 *>(d + span0 + j * 16u);
        const u32 f0 = ft_nl_flags(v.x), f1 = ft_nl_flags(v.y), f2 = ft_nl_flags(v.z), f3 = ft_nl_flags(v.w);
        u32 lo = __dp4a(f0, 0x08040201u, 0u);
        lo = __dp4a(f1, 0x80402010u, lo);
        u32 hi = __dp4a(f2, 0x08040201u, 0u);
        hi = __dp4a(f3, 0x80402010u, hi);
        m16[j] = (lo >> 7) | (hi << 1);
        const u32 bad = ft_not_acgt(v.x, f0) | ft_not_acgt(v.y, f1) | ft_not_acgt(v.z, f2) | ft_not_acgt(v.w, f3);
        if (bad == 0 && span0 + j * 16u + 16u <= lim) clean |= 1u << j;
      }
      m[0] = m16[0] | (m16[1] << 16);
      m[1] = m16[2] | (m16[3] << 16);
      m[2] = m16[4] | (m16[5] << 16);
      if (span0 + SPAN > lim) {  // padding behind the end of the file does not count
        const u32 valid = lim > span0 ? lim - span0 : 0u;
#pragma unroll
        for (u32 k2 = 0; k2 < 3; k2++) {
          const u32 lo = k2 * 32u;
          if (valid <= lo) m[k2] = 0;
          else if (valid < lo + 32u) m[k2] &= (1u << (valid - lo)) - 1u;
        }
      }
    }
    // newline count prefix, and the position (+1) of the last newline in front of every span
    const u32 cnt = (u32)(__popc(m[0]) + __popc(m[1]) + __popc(m[2]));
    u32 last = 0;  // region position + 1 of the span's last newline, 0 = none
    if (m[2]) last = span0 + 64u + 32u - (u32)__clz((int)m[2]);
    else if (m[1]) last = span0 + 32u + 32u - (u32)__clz((int)m[1]);
    else if (m[0]) last = span0 + 32u - (u32)__clz((int)m[0]);
    u32 inc = cnt, lmax = last;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const u32 y = __shfl_up_sync(0xffffffffu, inc, off);
      const u32 z = __shfl_up_sync(0xffffffffu, lmax, off);
      if ((int)lane >= off) { inc += y; lmax = lmax > z ? lmax : z; }
    }
    u32 prev_last = __shfl_up_sync(0xffffffffu, lmax, 1);  // last newline of the earlier lanes of this warp
    if (lane == 0) prev_last = 0;
    if (lane == 31) { s_wtot[warp] = inc; s_wlast[warp] = lmax; }
    if (tid == LBL) s_nl_lb = inc - cnt;
    __syncthreads();
    u32 nl_before = inc - cnt, nl_total = 0;
#pragma unroll
    for (u32 w = 0; w < NWARP; w++) {
      const u32 x = s_wtot[w];
      if (w < warp) { nl_before += x; const u32 z = s_wlast[w]; prev_last = prev_last > z ? prev_last : z; }
      nl_total += x;
    }
    const u32 nl_lookbehind = s_nl_lb;
    if (tid == 0) a.tile_nl[tile] = nl_total - nl_lookbehind;

    // ---- per newline of the owned range: header list, line-shape check; header bytes of the span (bit mask)
    u32 hm[3] = {0, 0, 0};
    if (tid >= LBL - 1u) {
      u32 xp1 = prev_last;  // start of the line that the next newline ends (position of the previous newline + 1)
      u32 seen = 0;
      bool in_hdr = prev_last > 0 && d[prev_last] == '>';  // the span starts inside a header line
      u32 pos = 0;                                         // first byte of the span not yet classified
      auto mark = [&](u32 from, u32 to) {                  // header bytes [from, to) of the span
#pragma unroll
        for (u32 q = 0; q < 3; q++) {
          const u32 lo = q * 32u, hi = lo + 32u;
          if (from < hi && to > lo) {
            const u32 f2 = from > lo ? from - lo : 0u, t2 = to < hi ? to - lo : 32u;
            hm[q] |= (t2 >= 32u ? 0xffffffffu : ((1u << t2) - 1u)) & ~((1u << f2) - 1u);
          }
        }
      };
#pragma unroll
      for (u32 k2 = 0; k2 < 3; k2++) {
        u32 mm = m[k2];
        while (mm) {
          const u32 b = (u32)__ffs((int)mm) - 1u;
          mm &= mm - 1u;
          const u32 x = span0 + k2 * 32u + b;  // region position of the newline
          seen++;
          const u32 h = x + 1u;
          const bool next_hdr = h < lim_a && d[h] == '>';
          if (in_hdr) mark(pos, k2 * 32u + b);
          in_hdr = next_hdr;
          pos = k2 * 32u + b + 1u;
          if (next_hdr && h >= LB && h < REGION) {  // a record starts inside the owned range
            const unsigned long long idx = atomicAdd((unsigned long long *)&a.st->counters[6], 1ull);
            if (idx < a.hdr_cap) {
              a.hdr_off[idx] = (u64)t0 + (h - LB);
              a.hdr_nl[idx] = (u64)(nl_before + seen - nl_lookbehind);  // newlines in [t0, h)
            }
          }
          if (x >= LB) {  // the line ending here: [xp1, x)
            const bool have_start = xp1 > 0 || t0 == 0;  // its start lies inside the region (or the file starts here)
            const u32 ls = xp1 > 0 ? xp1 : LB;
            const bool is_hdr = have_start && d[ls] == '>';
            const bool last_line = next_hdr || (eof && h >= lim);
            if (!have_start) {
              if (a.width) s_bad = 1;  // a line longer than the look-behind in a wrapped file
              else if (!last_line) s_bad = 1;
            } else if (!is_hdr) {
              const u32 len = x - ls;
              if (a.width ? (last_line ? len > a.width : len != a.width) : !last_line) s_bad = 1;
            }
          }
          xp1 = h;
        }
      }
      if (in_hdr) mark(pos, SPAN);
    }
    // chunks that are not plain A/C/G/T only because a header line runs through them: judge their sequence bytes alone
    if (tid >= LBL && span0 < lim && (hm[0] | hm[1] | hm[2])) {
#pragma unroll
      for (u32 j = 0; j < 6; j++) {
        const u32 hb = (hm[j >> 1] >> (16u * (j & 1u))) & 0xffffu;
        if (!((clean >> j) & 1u) && hb && span0 + j * 16u + 16u <= lim) {
          const uint4 v = *reinterpret_cast<const uint4 *>(d + span0 + j * 16u);
          const u32 wv[4] = {v.x, v.y, v.z, v.w};
          u32 bad = 0;
#pragma unroll
          for (u32 q = 0; q < 4; q++) {
            const u32 bm = ((((hb >> (4u * q)) & 0xfu) * 0x00204081u) & 0x01010101u) * 0xffu;  // 4 mask bits -> 4 mask bytes
            bad |= ft_not_acgt(wv[q], ft_nl_flags(wv[q])) & ~bm;
          }
          if (bad == 0) clean |= 1u << j;
        }
      }
    }
    if (tid >= LBL && span0 < lim) a.clean[(size_t)tile * (T / SPAN) + (tid - LBL)] = (u8)clean;
    if (tile == 0 && tid == 0 && n > 0 && a.in[0] != '>') s_bad = 1;  // the input must open with a record
    // an unterminated last line (no final newline): its shape is judged here
    if (eof && tid == 0 && n > 0 && d[lim - 1] != '\n') a.st->counters[7] = 1;  // the caller handles this rare case on the general path
    __syncthreads();
    if (tid == 0) {
      if (s_bad) { atomicAdd((unsigned long long *)&a.st->counters[0], 1ull); s_bad = 0; }
      const u32 tn = tile + NSTAGE * gridDim.x;
      if (tn < a.n_tiles) issue(tn, s);
    }
  }
}

u32 fasta_tile_tiles(u32 n) { return (n + ft::T - 1) / ft::T; }
u32 fasta_tile_bytes() { return ft::T; }
u32 fasta_tile_spans_per_tile() { return ft::T / ft::SPAN; }

void fasta_index_tile(FastaTileArgs a, int n_sm, cudaStream_t s) {
  a.n_tiles = fasta_tile_tiles(a.n);
  const size_t smem = sizeof(ft::Smem) + 16;
#ifndef BSK_EMU
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaFuncSetAttribute(k_fasta_index_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set[dev] = true;
  }
#endif
  u32 grid = (u32)n_sm * ft::CTAS;
  if (grid > a.n_tiles) grid = a.n_tiles;
  if (grid == 0) return;
  BSK_LAUNCH(k_fasta_index_tile, grid, ft::NT, smem, s, a);
}

}  // namespace k
}  // namespace bsk
