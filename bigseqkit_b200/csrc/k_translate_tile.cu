// k_translate_tile.cu -- translate on wrapped FASTA read IN PLACE, every (record, frame) element formatted straight
// into the output (BASELINE configs[4]: translate --frame 6 on CDS-like records wrapped at 60).
//
//   Translate.Call    bigseqkit-lib/translate.go:66-145: per record and frame Seq.Translate(table, frame, ...), header
//                     ">Name" (:133-137), WrapByteSlice(protein, LineWidth) (:138); SURVEY Q3: one element per frame
//   Seq.Translate     bio v0.7.0 (restated in oracle/bsk_oracle.c): frame < 0 -> reverse complement first; codons from
//                     offset |frame| - 1; ambiguous codons: every expansion agrees -> that amino acid, else 'X'
//
// The general path squeezes the newlines out of the sequences, translates into an arena and formats that arena in a
// third pass.  Here the kernel is driven by the OUTPUT: a CTA formats 4096 output bytes in shared memory; its work is
// cut into (element, 16-byte window) pieces and one header item per element, and for a piece it
//   * (the usual case) fetches the <= 48 bases of its <= 16 amino acids from the raw input with four
//     aligned 16-byte loads -- base b of a record sits at seq_start + b + b / width, which k_fasta_index_tile has
//     verified for the whole block -- packs them to 2 bits each (the span flags of the index pass say that they are
//     plain A/C/G/T), removes the slot of the one line break that can fall inside, and looks every codon up in a
//     64-entry table (a second table holds the reverse-complement strand: no complemented copy of the sequence);
//     the wrap newline of the output line is inserted with funnel shifts; one aligned 16-byte store;
//   * anywhere else (header bytes, element boundaries, ambiguous or lower-case bases, gaps, narrow wrapping): byte by
//     byte through the IUPAC tables of the general path, same rules, same error reporting.
// HBM traffic = N read (the frames of a record are formatted next to each other, so their re-reads hit L2) + output
// written once.
#include "kernels.h"

namespace bsk {
namespace k {

struct TrFrames { int f[8]; };

__global__ void k_translate_sizes(const u32 *__restrict__ name_len, const u32 *__restrict__ seq_len, u32 n_rec, u32 nf, TrFrames fr,
                                  u32 width_out, u32 *__restrict__ sizes, DevStatus *st) {
  const u32 e = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 n_el = n_rec * nf;
  if (e > n_el) return;
  if (e == n_el) { sizes[e] = 0; return; }
  const u32 r = e / nf, fi = e - r * nf;
  const u32 l = seq_len[r];
  u32 plen = 0;
  if (l < 3) {
    atomicMin((unsigned long long *)&st->err, ((unsigned long long)r << 4) | EK_TOO_SHORT);
  } else {
    const int f = fr.f[fi];
    plen = (l - (u32)((f < 0 ? -f : f) - 1)) / 3;
  }
  const u32 wrapl = plen ? plen + (width_out ? (plen - 1) / width_out : 0u) : 0u;
  sizes[e] = 1u + name_len[r] + 1u + wrapl + 1u;  // '>' header '\n' wrapped protein, FileStore's '\n'
}

// four bases (one per byte of x) -> 8 bits, first base in the low bits; A 0, C 1, T/U 2, G 3
__device__ __forceinline__ u32 tt_pack4(u32 x) { return ((((x >> 1) & 0x03030303u) * 0x01041040u) >> 24); }
// bits [0, nbits) of a 32-bit word set (nbits may exceed 32 or be negative)
__device__ __forceinline__ u32 tt_lowmask(int nbits) {
#ifdef BSK_EMU
  return nbits >= 32 ? 0xffffffffu : (nbits <= 0 ? 0u : ((1u << nbits) - 1u));
#else
  return __funnelshift_lc(0xffffffffu, 0u, (u32)max(nbits, 0));  // the shift clamps at 32
#endif
}
// q / d for a block-wide constant d <= 1001 and q < 2^30: one wide multiply (m = ceil(2^40 / d))
__device__ __forceinline__ u32 tt_div(u32 q, unsigned long long m) { return (u32)(((unsigned long long)q * m) >> 40); }

namespace tt {
constexpr u32 NT = 256, TILE = 4096, ICAP = 1664;  // ICAP: one item per window + one per element that starts in the tile (>= 3 bytes each)
struct El {  // one (record, frame) element
  u64 eo;    // offset of its '>' in the output
  u32 r, H, l, s0, start, plen, wrapl, name_off;
  int f;
};
}  // namespace tt

__device__ __forceinline__ void tt_load_el(const TranslateTileArgs &a, u32 e, tt::El &E) {
  E.eo = a.out_off[e];
  E.r = e / a.nf;
  E.f = a.frames[e - E.r * a.nf];
  E.H = a.name_len[E.r];
  E.l = a.seq_len[E.r];
  E.s0 = a.seq_start[E.r];
  E.name_off = a.name_off[E.r];
  E.start = (u32)((E.f < 0 ? -E.f : E.f) - 1);
  E.plen = E.l >= 3 ? (E.l - E.start) / 3 : 0u;
  E.wrapl = E.plen ? E.plen + (a.width_out ? (E.plen - 1) / a.width_out : 0u) : 0u;
}

// amino acid j of the element, byte by byte through the IUPAC tables (ambiguous / lower-case bases, gaps, errors)
// (rare: the tables are read from global memory, L1 / L2 hits, instead of taking 4.5 KiB of every CTA's shared memory)
__device__ __forceinline__ u32 tt_code(const u8 *tab, u8 b) {
  const u32 x = tab[b];
  return x == 16u ? 0x10u : (x == 0u ? 0x20u : x);  // IUPAC mask 1..15; gap 0x10; not a nucleotide 0x20
}
__device__ __forceinline__ u8 tt_aa_careful(const TranslateTileArgs &a, const tt::El &E, u32 j) {
  const u32 Wi = a.width_in, i = E.start + 3u * j, l = E.l;
  u32 b0, b1, b2;  // positions on the '+' strand
  if (E.f > 0) { b0 = i; b1 = i + 1u; b2 = i + 2u; }
  else { b0 = l - 1u - i; b1 = l - 2u - i; b2 = l - 3u - i; }
  const u8 *tab = E.f > 0 ? a.code_fwd : a.code_rev;
  const u32 c0 = tt_code(tab, a.in[E.s0 + b0 + (Wi ? b0 / Wi : 0u)]);
  const u32 c1 = tt_code(tab, a.in[E.s0 + b1 + (Wi ? b1 / Wi : 0u)]);
  const u32 c2 = tt_code(tab, a.in[E.s0 + b2 + (Wi ? b2 / Wi : 0u)]);
  u8 c;
  bool init = false;
  if (c0 == 0x10 && c1 == 0x10 && c2 == 0x10) c = '-';
  else if ((c0 | c1 | c2) & 0x30u) {
    c = 'X';
    if (!a.allow_unknown) atomicMin((unsigned long long *)&a.st->err, ((unsigned long long)E.r << 4) | EK_UNKNOWN_CODON);
  } else {
    const u8 x = a.lut[(c0 << 8) | (c1 << 4) | c2];
    c = x & 0x7f;
    init = (x & 0x80) != 0;
  }
  if (a.init_m && j == 0 && init) c = 'M';
  if (a.clean_stop && c == '*') c = 'X';
  return c;
}

// amino acids j0 .. j0 + naa - 1 (naa <= 16) of the element from plain A/C/G/T bases read in place: v[] receives them
// packed 4 per word.  false when the bases are not plain (span flags), the input is wrapped too narrowly or the 64-byte
// load window leaves the buffer: the caller then takes the careful path.
__device__ __forceinline__ bool tt_aa_fast(const TranslateTileArgs &a, const tt::El &E, u32 j0, u32 naa, const u8 *s_aa, u32 v[4]) {
  const u32 Wi = a.width_in;
  if ((Wi && Wi < 48u) || E.l >= (1u << 30)) return false;
  const u32 nb = 3u * naa, i0 = E.start + 3u * j0;
  const u32 fb0 = E.f > 0 ? i0 : E.l - i0 - nb;  // first base of the window on the '+' strand
  u32 line_in = 0, sb = 64;                      // sb: bases of the window in front of the line break inside it
  if (Wi) {
    line_in = tt_div(fb0, a.magic_in);
    const u32 to_break = (line_in + 1u) * Wi - fb0;
    if (to_break < nb) sb = to_break;
  }
  const u32 raw0 = E.s0 + fb0 + line_in;
  const u32 a0 = raw0 & ~15u, off = raw0 & 15u;
  if (a0 + 64u > a.n) return false;
  {  // every chunk the window touches must hold plain A/C/G/T sequence bytes
    const u32 c0 = a0 >> 4, nch = (off + nb + (sb < 64u ? 1u : 0u) + 15u) >> 4;
    const u32 gs = c0 / 6u, jj = c0 - gs * 6u;
    const u32 bits = ((u32)a.clean[gs] | ((u32)a.clean[gs + 1] << 6)) >> jj;
    const u32 need = (1u << nch) - 1u;
    if ((bits & need) != need) return false;
  }
  const uint4 *vp = reinterpret_cast<const uint4 *>(a.in + a0);
  const uint4 q0 = vp[0], q1 = vp[1], q2 = vp[2], q3 = vp[3];
  u32 P[5];
  P[0] = tt_pack4(q0.x) | (tt_pack4(q0.y) << 8) | (tt_pack4(q0.z) << 16) | (tt_pack4(q0.w) << 24);
  P[1] = tt_pack4(q1.x) | (tt_pack4(q1.y) << 8) | (tt_pack4(q1.z) << 16) | (tt_pack4(q1.w) << 24);
  P[2] = tt_pack4(q2.x) | (tt_pack4(q2.y) << 8) | (tt_pack4(q2.z) << 16) | (tt_pack4(q2.w) << 24);
  P[3] = tt_pack4(q3.x) | (tt_pack4(q3.y) << 8) | (tt_pack4(q3.z) << 16) | (tt_pack4(q3.w) << 24);
  P[4] = 0;
  u32 S[4];  // stream of the window: position k (2 bits) = byte raw0 + k
#pragma unroll
  for (int q = 0; q < 4; q++) S[q] = __funnelshift_r(P[q], P[q + 1], 2u * off);
  u32 R[3];  // ... without the slot of the line break
  {
    const int cut = (int)(2u * sb);
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const u32 keep = tt_lowmask(cut - 32 * q);
      R[q] = (S[q] & keep) | (__funnelshift_r(S[q], S[q + 1], 2u) & ~keep);
    }
  }
  // The '-' strand reads the window downwards: move it up so that it ends at position 48 and reverse the order of the
  // 48 two-bit groups -- codon t then sits at bits 6t .. 6t+5 with its FIRST base lowest, as on the '+' strand, and both
  // strands share the look-up loop (s_aa[64 ..] is the complement-strand table under that index order).
  u32 toff = 0;
  if (E.f < 0) {
    const u32 sh = 6u * (16u - naa);
    u32 r0 = R[0], r1 = R[1], r2 = R[2];
    if (sh) {
      const u32 ws = sh >> 5, bs = sh & 31u;
      const u32 x0 = r0, x1 = r1, x2 = r2;
      if (ws == 0) { r2 = __funnelshift_l(x1, x2, bs); r1 = __funnelshift_l(x0, x1, bs); r0 = x0 << bs; }
      else if (ws == 1) { r2 = __funnelshift_l(x0, x1, bs); r1 = x0 << bs; r0 = 0; }
      else { r2 = x0 << bs; r1 = 0; r0 = 0; }
    }
    r0 = __brev(r0); r1 = __brev(r1); r2 = __brev(r2);  // bit reversal, then the two bits of every group back in order
    R[0] = ((r2 >> 1) & 0x55555555u) | ((r2 & 0x55555555u) << 1);
    R[1] = ((r1 >> 1) & 0x55555555u) | ((r1 & 0x55555555u) << 1);
    R[2] = ((r0 >> 1) & 0x55555555u) | ((r0 & 0x55555555u) << 1);
    toff = 64u;
  }
  u32 aa[16];
#pragma unroll
  for (int t = 0; t < 16; t++) {
    const int bit = 6 * t;
    const u32 idx = (bit + 6 <= 32 * (bit / 32 + 1) ? (R[bit / 32] >> (bit & 31)) : __funnelshift_r(R[bit / 32], R[bit / 32 + 1], bit & 31)) & 63u;
    aa[t] = s_aa[idx | toff];
    if (t == 0 && a.init_m && j0 == 0) {
      // start codons: bit per index of the 64-entry tables as the host built them (complement strand: third base lowest)
      const u32 io = toff ? (((idx & 3u) << 4) | (idx & 0xcu) | (idx >> 4)) : idx;
      if (((toff ? a.start_rev : a.start_fwd) >> io) & 1ull) aa[0] = 'M';
    }
  }
#pragma unroll
  for (int q = 0; q < 4; q++) v[q] = aa[4 * q] | (aa[4 * q + 1] << 8) | (aa[4 * q + 2] << 16) | (aa[4 * q + 3] << 24);
  return true;
}

// the bytes of 16-byte window w of the tile that belong to the wrapped protein of element E -> s_out
__device__ __forceinline__ void tt_piece(const TranslateTileArgs &a, const tt::El &E, u32 w, u64 cta0, u8 *s_out, const u8 *s_aa) {
  const u32 Wo = a.width_out;
  const u64 wa = cta0 + 16ull * w, pa = E.eo + E.H + 2u, pb = pa + E.wrapl;
  if (pa >= wa + 16u || pb <= wa) return;  // no protein byte of this element in the window
  const u32 k0 = pa > wa ? (u32)(pa - wa) : 0u;
  const u32 k1 = pb < wa + 16u ? (u32)(pb - wa) : 16u;
  const u32 q0 = (u32)(wa + k0 - pa);  // offset of the piece in the wrapped protein
  u32 j0 = q0, nlpos = 16;             // first amino acid; byte of the window that is the wrap '\n' (16 = none)
  bool simple = true;                  // at most one wrap newline inside the piece
  if (Wo) {
    const u32 line = tt_div(q0, a.magic_out1), col = q0 - line * (Wo + 1u);
    j0 = line * Wo + col;
    if (k0 + (Wo - col) < k1) nlpos = k0 + (Wo - col);
    simple = Wo >= 16u;
  }
  const u32 naa = (k1 - k0) - (nlpos < 16u ? 1u : 0u);
  u32 v[4];
  if (simple && naa && E.wrapl < (1u << 30) && tt_aa_fast(a, E, j0, naa, s_aa, v)) {
    if (nlpos < 16u) {  // the output line ends inside the piece: shift the tail up by one byte, put the '\n' in
      const int cut = (int)(8u * (nlpos - k0));
      u32 up[4];
      up[0] = v[0] << 8;
#pragma unroll
      for (int q = 1; q < 4; q++) up[q] = __funnelshift_l(v[q - 1], v[q], 8u);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const u32 keep = tt_lowmask(cut - 32 * q), nl = tt_lowmask(cut + 8 - 32 * q) & ~keep;
        v[q] = (v[q] & keep) | (0x0a0a0a0au & nl) | (up[q] & ~(keep | nl));
      }
    }
    if (k0 == 0 && k1 == 16u) {
      *reinterpret_cast<uint4 *>(s_out + 16u * w) = make_uint4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (u32 k2 = 0; k2 < 16u; k2++)
        if (k2 < k1 - k0) s_out[16u * w + k0 + k2] = (u8)(v[k2 >> 2] >> (8u * (k2 & 3u)));
    }
  } else {  // careful path: byte by byte
    for (u32 k2 = k0; k2 < k1; k2++) {
      const u32 q = q0 + (k2 - k0);
      u32 j = q;
      bool is_nl = false;
      if (Wo) {
        const u32 line = q / (Wo + 1u), col = q - line * (Wo + 1u);
        is_nl = col == Wo;
        j = line * Wo + col;
      }
      s_out[16u * w + k2] = is_nl ? (u8)'\n' : tt_aa_careful(a, E, j);
    }
  }
}

// window w of the tile lies wholly inside the wrapped protein of E: 16 bytes from plain bases, at most one wrap newline.
// false (nothing written) when the window needs the careful path -- the caller leaves it to tt_piece.
__device__ __forceinline__ bool tt_full(const TranslateTileArgs &a, const tt::El &E, u32 w, u64 cta0, u8 *s_out, const u8 *s_aa) {
  const u32 Wo = a.width_out;
  if ((Wo && Wo < 16u) || E.wrapl >= (1u << 30)) return false;
  const u32 q0 = (u32)(cta0 + 16ull * w - (E.eo + E.H + 2u));  // offset of the window in the wrapped protein
  u32 j0 = q0, nlpos = 16;
  if (Wo) {
    const u32 line = tt_div(q0, a.magic_out1), col = q0 - line * (Wo + 1u);
    j0 = line * Wo + col;
    if (Wo - col < 16u) nlpos = Wo - col;
  }
  u32 v[4];
  if (!tt_aa_fast(a, E, j0, nlpos < 16u ? 15u : 16u, s_aa, v)) return false;
  if (nlpos < 16u) {  // the output line ends inside the window: shift the tail up by one byte, put the '\n' in
    const int cut = (int)(8u * nlpos);
    u32 up[4];
    up[0] = v[0] << 8;
#pragma unroll
    for (int q = 1; q < 4; q++) up[q] = __funnelshift_l(v[q - 1], v[q], 8u);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const u32 keep = tt_lowmask(cut - 32 * q), nl = tt_lowmask(cut + 8 - 32 * q) & ~keep;
      v[q] = (v[q] & keep) | (0x0a0a0a0au & nl) | (up[q] & ~(keep | nl));
    }
  }
  *reinterpret_cast<uint4 *>(s_out + 16u * w) = make_uint4(v[0], v[1], v[2], v[3]);
  return true;
}

// A CTA formats 4096 output bytes in shared memory and writes them with aligned 16-byte stores: warps copy the headers
// of the elements that touch the tile, every thread fills the protein bytes of one 16-byte window for the element the
// window starts in, and the few windows that reach into a further element's protein leave an item in a short list
// that is worked off round-robin afterwards (no divergent second pass of a whole warp).
__global__ void __launch_bounds__(tt::NT, 8) k_translate_tile(TranslateTileArgs a) {
  using namespace tt;
  constexpr u32 ECAP = 64;  // elements of the tile whose geometry is cached in shared memory
  __align__(16) __shared__ u8 s_out[TILE];
  __shared__ u8 s_aa[128];  // plain-codon tables: '+' strand, then the complement strand with the first base lowest
  __shared__ El s_el[ECAP];
  __shared__ u32 s_over[ICAP];
  __shared__ u32 s_cnt, s_n;
  const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  {
    if (tid < 64) { s_aa[tid] = a.aa_fwd[tid]; s_aa[64u + (((tid & 3u) << 4) | (tid & 0xcu) | (tid >> 4))] = a.aa_rev[tid]; }
    if (tid == 0) { s_n = 0; s_cnt = 0xffffffffu; }
  }
  const u32 n_el = a.n_rec * a.nf;
  const u64 cta0 = (u64)blockIdx.x * TILE;
  const u64 cta1 = cta0 + TILE < a.total ? cta0 + TILE : a.total;  // end of this tile's bytes
  const u32 e_lo = a.tile_first[blockIdx.x];  // last element that starts at or before the tile (k_translate_first)
  __syncthreads();
  // elements e_lo .. e_lo + cnt - 1 touch the tile: cnt = first k >= 1 with out_off[e_lo + k] >= cta1
  for (u32 k2 = tid + 1;; k2 += NT) {
    const u32 e = e_lo + k2 < n_el ? e_lo + k2 : n_el;
    const bool past = a.out_off[e] >= cta1;
    if (past) atomicMin(&s_cnt, k2);
    if (__syncthreads_or(past ? 1 : 0)) break;
  }
  const u32 cnt = s_cnt;
  if (tid < cnt && tid < ECAP) tt_load_el(a, e_lo + tid, s_el[tid]);
  __syncthreads();

  // ---- '>' + header + '\n' in front of every protein, '\n' behind it (clipped to the tile): one warp per element
  for (u32 k2 = warp; k2 < cnt; k2 += NT / 32) {
    El E;
    if (k2 < ECAP) E = s_el[k2];
    else tt_load_el(a, e_lo + k2, E);
    const u64 hb = E.eo, he = E.eo + E.H + 2u;  // [hb, he)
    const u64 from = hb > cta0 ? hb : cta0, to = he < cta1 ? he : cta1;
    for (u64 p = from + lane; p < to; p += 32) {
      const u32 relb = (u32)(p - hb);
      s_out[p - cta0] = relb == 0 ? (u8)'>' : (relb <= E.H ? a.in[E.name_off + relb - 1u] : (u8)'\n');
    }
    const u64 term = he + E.wrapl;
    if (lane == 0 && term >= cta0 && term < cta1) s_out[term - cta0] = '\n';
  }
  // ---- proteins: the window's own element here, further elements that start inside the window through the list
  {
    const u64 wa = cta0 + 16ull * tid;
    if (wa < cta1) {
      u32 lo = 0, hi = cnt;  // last k with out_off[e_lo + k] <= wa
      while (hi - lo > 1) {
        const u32 mid = lo + ((hi - lo) >> 1);
        const u64 om = mid < ECAP ? s_el[mid].eo : a.out_off[e_lo + mid];
        if (om <= wa) lo = mid;
        else hi = mid;
      }
      El E;
      if (lo < ECAP) E = s_el[lo];
      else tt_load_el(a, e_lo + lo, E);
      // a window inside one protein is done here; windows at the ends of a protein (one or two lanes of every warp) and
      // the ones that need the careful path join the list, so that no warp runs the general piece code for one lane
      const u64 pa = E.eo + E.H + 2u, pb = pa + E.wrapl;
      bool done = pa >= wa + 16u || pb <= wa;  // no protein byte of the element in the window
      if (!done && pa <= wa && wa + 16u <= pb) done = tt_full(a, E, tid, cta0, s_out, s_aa);
      if (!done) {
        const u32 slot = atomicAdd(&s_n, 1u);
        if (slot < ICAP) s_over[slot] = (lo << 8) | tid;
      }
      for (u32 k2 = lo + 1; k2 < cnt; k2++) {
        const u64 on = k2 < ECAP ? s_el[k2].eo : a.out_off[e_lo + k2];
        if (on >= wa + 16u) break;
        const u32 slot = atomicAdd(&s_n, 1u);
        if (slot < ICAP) s_over[slot] = (k2 << 8) | tid;
      }
    }
  }
  __syncthreads();
  const u32 n_over = s_n < ICAP ? s_n : ICAP;  // cannot overflow: one item per window + one per element that starts inside the tile
  for (u32 it = tid; it < n_over; it += NT) {
    const u32 item = s_over[it], k2 = item >> 8;
    El E;
    if (k2 < ECAP) E = s_el[k2];
    else tt_load_el(a, e_lo + k2, E);
    tt_piece(a, E, item & 0xffu, cta0, s_out, s_aa);
  }
  __syncthreads();
  const u64 o = cta0 + 16ull * tid;
  if (o + 16ull <= a.total) {
    *reinterpret_cast<uint4 *>(a.out + o) = *reinterpret_cast<const uint4 *>(s_out + 16u * tid);
  } else {
    for (u32 k2 = 0; o + k2 < a.total; k2++) a.out[o + k2] = s_out[16u * tid + k2];
  }
}

// first element of every 4096-byte output tile: one thread per tile, binary search over the element offsets
__global__ void k_translate_first(const u64 *__restrict__ out_off, u32 n_el, u64 total, u32 *__restrict__ tile_first) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 o = t * tt::TILE;
  if (o >= total) return;
  u32 lo = 0, hi = n_el;  // last e with out_off[e] <= o
  while (hi - lo > 1) {
    const u32 mid = lo + ((hi - lo) >> 1);
    if (out_off[mid] <= o) lo = mid;
    else hi = mid;
  }
  tile_first[t] = lo;
}

void translate_sizes(const u32 *name_len, const u32 *seq_len, u32 n_rec, u32 nf, const int *frames, u32 width_out, u32 *sizes,
                     DevStatus *st, cudaStream_t s) {
  TrFrames fr;
  for (u32 i = 0; i < 8; i++) fr.f[i] = i < nf ? frames[i] : 1;
  const u32 n_el = n_rec * nf;
  BSK_LAUNCH_FLAT(k_translate_sizes, (n_el + 1 + 255) / 256, 256, 0, s, name_len, seq_len, n_rec, nf, fr, width_out, sizes, st);
}

u32 translate_tile_tiles(u64 total) { return (u32)((total + tt::TILE - 1) / tt::TILE); }

void translate_tile(const TranslateTileArgs &a, cudaStream_t s) {
  if (!a.total) return;
  const u32 nt = translate_tile_tiles(a.total);
  BSK_LAUNCH_FLAT(k_translate_first, (nt + 255) / 256, 256, 0, s, a.out_off, a.n_rec * a.nf, a.total, a.tile_first);
  BSK_LAUNCH(k_translate_tile, (u32)((a.total + tt::TILE - 1) / tt::TILE), tt::NT, 0, s, a);
}

}  // namespace k
}  // namespace bsk
