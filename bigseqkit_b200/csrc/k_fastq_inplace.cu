// k_fastq_inplace.cu -- `seq` on 4-line FASTQ when the output record has the layout of the input record
// (full record printed, no filter, bare "+" line): BASELINE configs[1], `seq --reverse --complement`.
//
//   PlainFile split + ReadFixer     bigseqkit/helper.go:148-178, bigseqkit-lib/helper.go:41-66
//   SeqParser.Read                  bigseqkit-lib/helper.go:219-325
//   SeqTransform.Call               bigseqkit-lib/seq.go:81-269  (reverse :188-190, complement :191-196,
//                                   dna2rna/rna2dna/case :199-239, FASTQ is never wrapped :123)
//   FileStore framing               bigseqkit-lib/helper.go:441-451
//
// In that mode byte i of the output stream depends only on the record that covers byte i of the
// input, and sits at the same offset.  So there is no inter-CTA dependency at all:
//
//   * persistent CTAs (512 threads, 4 per SM = 64 warps at 32 registers; other shapes stay selectable for A/B runs
//     with BSK_FQ_SHAPE) walk 20 KiB tiles (+ 4 KiB halo) with a 2-stage ring filled by 1-D TMA bulk loads;
//   * a CTA finds the newlines of its region (16-byte shared-memory loads, SWAR zero-byte test, IDP.4A mask
//     packing) and turns them into the list of line starts;
//   * the records a tile owns (the ones that START inside it) are a chain of 4-line groups behind the first
//     record line, so record r sits at lines kmin + 4r: their number falls out of one counting barrier, and the
//     lanes that rewrite a record check it against the grammar of SeqParser.Read first;
//   * the sequence (reverse + 256-entry byte map) and the quality (reverse) of every owned record are
//     rewritten IN PLACE in the stage buffer by 8 lanes, each producing consecutive 32-bit words with PRMT from
//     a descending run of source words; the ragged first / last word of a segment is blended with the old word;
//   * the owned byte range leaves through one TMA bulk store (+ <= 15 ragged bytes at each end);
//   * record starts go to a per-tile slot list; a tiny second kernel turns the per-tile counts (scanned
//     with cub) into the global element-offset array.
//
// HBM traffic = N read + N written (+ the 4 KiB halo re-read, served by L2, and 2 B per record of slots).
// Anything outside the grammar (multi-line records, "+name" lines, missing marker, blank lines, records
// longer than the halo, unmatched lengths) raises a flag and the caller re-runs the block on the general
// path, which also produces the reference's error text.
#include <cstdlib>

#include "kernels.h"
#include "tma.cuh"

namespace bsk {
namespace k {

namespace fq {
constexpr u32 H = 4096;        // halo bytes (longest record the kernel accepts, roughly)
constexpr u32 PRE = 16;        // look-behind bytes in front of the tile
constexpr u32 RCAP = 1024;     // owned records per tile (slot stride)
// CTA shape: NT threads, every lane scans CPL consecutive 16-byte chunks, so tile + halo = NT * CPL * 16 bytes.
constexpr u32 CPL = 3, NSTAGE = 2;
template <u32 NT_, u32 CTAS_, u32 LCAP_>
struct Cfg {
  static constexpr u32 NT = NT_, CTAS = CTAS_, LCAP = LCAP_;  // threads, CTAs per SM, line starts per region
  static constexpr u32 NWARP = NT / 32;
  static constexpr u32 T = NT * CPL * 16 - H;  // tile bytes
  static constexpr u32 STAGE = PRE + T + H + 16;
  static_assert(T % 16 == 0 && STAGE % 16 == 0 && T + H < 65536 && NWARP <= 32 && NT <= RCAP, "tile geometry");
  struct Smem {
    u8 lut[256];                 // first, on a 256-byte boundary: lut4() forms addresses with PRMT instead of adds
    u8 in[NSTAGE][STAGE];
    u64 full[NSTAGE];
    u16 ls[LCAP + 8];            // line starts, ls[0] = 0
    u32 wtot[NWARP];
    u32 bad, rescan, kmin;
    u32 bad_rec;                 // a record failed the grammar check of the transform pass (read by warp 0 behind it)
  };
};
typedef Cfg<512, 4, 2048> CfgA;  // 64 warps / SM at 32 registers, 20 KiB tiles: the default
typedef Cfg<512, 3, 3072> CfgB;  // 48 warps / SM at 40 registers
typedef Cfg<256, 8, 1024> CfgC;  // 8 KiB tiles, 8 independent CTAs / SM
typedef Cfg<384, 5, 1536> CfgD;  // 14 KiB tiles, 5 CTAs / SM
typedef Cfg<1024, 2, 4096> CfgE; // 44 KiB tiles, 2 CTAs / SM of 32 warps
}  // namespace fq

struct FqInplaceArgs {
  const u8 *in;
  u32 n;
  u8 *out;
  const u8 *lut;
  u32 *tile_cnt;   // [n_tiles] owned records per tile
  u16 *slots;      // [n_tiles * RCAP] record starts relative to the tile
  DevStatus *st;   // counters[0] = fallback flag
  u32 n_tiles;
  int reverse, use_lut;
  int group;       // lanes per record in the transform: 8 (reads), 32
  int wpl;         // group == 8: words per lane (5 covers segments up to 157 B, else 8)
  u32 scan_halo;   // bytes of the halo the newline scan covers on its first attempt (multiple of 16, <= H)
};

// flags (0x80 per byte) of the bytes of w that equal '\n'; exact for every byte value.  (x | 0x80) - 1 keeps bit 7 of a
// byte unless its low seven bits are zero, and never borrows across bytes; bit 7 of w ^ '\n' is bit 7 of w.
__device__ __forceinline__ u32 nl_flags(u32 w) {
  const u32 t = ((w ^ 0x0a0a0a0au) | 0x80808080u) - 0x01010101u;
  return ~(t | w) & 0x80808080u;
}

// four table look-ups.  The table sits on a 256-byte boundary of shared memory, so the address of entry b is the table
// address with its low byte replaced by b: one PRMT per byte instead of extract + add.
#ifndef BSK_EMU
struct Lut {
  u32 sa;  // shared-window address of the table
  __device__ __forceinline__ explicit Lut(const u8 *lut) : sa(tma::smem_addr(lut)) {}
  __device__ __forceinline__ u32 at(u32 addr) const {
    u32 r;
    asm("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(addr));  // the table is read-only while the kernel runs
    return r;
  }
  __device__ __forceinline__ u32 one(u32 b) const { return at(__byte_perm(b, sa, 0x7650)); }
  __device__ __forceinline__ u32 four(u32 v) const {
    const u32 x0 = at(__byte_perm(v, sa, 0x7650)), x1 = at(__byte_perm(v, sa, 0x7651));
    const u32 x2 = at(__byte_perm(v, sa, 0x7652)), x3 = at(__byte_perm(v, sa, 0x7653));
    return x0 | (x1 << 8) | (x2 << 16) | (x3 << 24);
  }
};
#else
struct Lut {
  const u8 *t;
  explicit Lut(const u8 *lut) : t(lut) {}
  u32 one(u32 b) const { return t[b & 0xffu]; }
  u32 four(u32 v) const {
    return (u32)t[v & 0xffu] | ((u32)t[(v >> 8) & 0xffu] << 8) | ((u32)t[(v >> 16) & 0xffu] << 16) | ((u32)t[v >> 24] << 24);
  }
};
#endif

// The 32-bit words that cover the byte range [a, e) of the stage buffer, rewritten by a group of lanes: lane gl
// produces the WPL consecutive words behind word gl * WPL of the range.  Reversed, destination byte p takes source
// byte a + e - 1 - p, so a destination word is one PRMT of two neighbouring source words, and consecutive
// destination words walk down the source words: WPL + 1 loads per lane.  The first and the last word of the range
// are blended with the bytes they hold outside [a, e), so there is no byte path.
template <bool REV, bool LUT, u32 WPL>
__device__ __forceinline__ void seg_read(const u32 *w32, const Lut &lut, u32 a, u32 e, u32 gl, u32 (&v)[WPL], int &wi, int &nv) {
  const int wi0 = (int)(a >> 2);
  const int nW = e > a ? (int)((e + 3u) >> 2) - wi0 : 0;
  const int i0 = (int)(gl * WPL);
  const int rem = nW - i0;  // words of the range from this lane's first one on
  nv = rem < (int)WPL ? rem : (int)WPL;
  wi = wi0 + i0;
#pragma unroll
  for (int j = 0; j < (int)WPL; j++) v[j] = 0;
  if (nv <= 0) return;
  // the last word of the range (held by the lane with rem <= WPL) keeps its bytes from e on
  const bool last = rem <= (int)WPL && (e & 3u);
  const int jl = last ? nv - 1 : -1;
  const u32 ml = 0xffffffffu >> (8u * (4u - (e & 3u)));
  const u32 oldl = last ? w32[wi + nv - 1] & ~ml : 0u;
  int qb = 0;
  u32 sel = 0, hi = 0;
  if (REV) {
    const int v3 = (int)(a + e) - 4 - 4 * wi0;  // lowest source byte of destination word wi0
    sel = 0x0123u + (u32)(v3 & 3) * 0x1111u;    // PRMT selector {sh+3, sh+2, sh+1, sh}
    qb = (v3 >> 2) - i0;
    hi = w32[qb + 1];
  }
#pragma unroll
  for (int j = 0; j < (int)WPL; j++)
    if (j < nv) {
      u32 x;
      if (REV) {
        const u32 lo = w32[qb - j];
        x = __byte_perm(lo, hi, sel);
        hi = lo;
      } else {
        x = w32[wi + j];
      }
      if (LUT) x = lut.four(x);
      if (j == jl) x = (x & ml) | oldl;
      v[j] = x;
    }
  if (i0 == 0 && (a & 3u)) {  // the first word keeps its bytes in front of a
    const u32 m = 0xffffffffu << (8u * (a & 3u));
    v[0] = (v[0] & m) | (w32[wi] & ~m);
  }
}
template <u32 WPL>
__device__ __forceinline__ void seg_write(u32 *w32, const u32 (&v)[WPL], int wi, int nv) {
#pragma unroll
  for (int j = 0; j < (int)WPL; j++)
    if (j < nv) w32[wi + j] = v[j];
}

// The same for a range of G*WPL/2 .. G*WPL words (reads of one length: every record of the usual input): the lower
// half of the lanes counts its words from the first word of the range, the upper half from the last one.  The two
// runs meet or overlap in the middle, where both produce the same words, so no word needs a range test and the two
// blended words sit at fixed places (lane 0 word 0, lane G-1 word WPL-1).
template <bool REV, bool LUT, u32 G, u32 WPL>
__device__ __forceinline__ void seg_read_full(const u32 *w32, const Lut &lut, u32 a, u32 e, u32 gl, u32 (&v)[WPL], int &wi) {
  const int wi0 = (int)(a >> 2);
  const int nW = (int)((e + 3u) >> 2) - wi0;
  const int i0 = gl < G / 2 ? (int)(gl * WPL) : nW - (int)((G - gl) * WPL);
  wi = wi0 + i0;
  if (REV) {
    const int v3 = (int)(a + e) - 4 - 4 * wi0;          // lowest source byte of destination word wi0
    const u32 sel = 0x0123u + (u32)(v3 & 3) * 0x1111u;  // PRMT selector {sh+3, sh+2, sh+1, sh}
    const int qb = (v3 >> 2) - i0;
    u32 hi = w32[qb + 1];
#pragma unroll
    for (int j = 0; j < (int)WPL; j++) {
      const u32 lo = w32[qb - j];
      v[j] = __byte_perm(lo, hi, sel);
      hi = lo;
    }
  } else {
#pragma unroll
    for (int j = 0; j < (int)WPL; j++) v[j] = w32[wi + j];
  }
  if (LUT) {
#pragma unroll
    for (int j = 0; j < (int)WPL; j++) v[j] = lut.four(v[j]);
  }
  if (gl == 0 && (a & 3u)) {  // the first word keeps its bytes in front of a
    const u32 m = 0xffffffffu << (8u * (a & 3u));
    v[0] = (v[0] & m) | (w32[wi] & ~m);
  }
  if (gl == G - 1 && (e & 3u)) {  // the last word keeps its bytes from e on
    const u32 m = 0xffffffffu >> (8u * (4u - (e & 3u)));
    v[WPL - 1] = (v[WPL - 1] & m) | (w32[wi + (int)WPL - 1] & ~m);
  }
}

// In-place rewrite of one record by a group of G lanes: sequence [so, so+sl) reversed (REV) and mapped through
// lut (LUT), quality [qo, qo+sl) reversed (REV).  Registers hold a segment until every lane has read it, so the
// rewrite is hazard-free; the two segments share no word (at least "\n+\n" lies between them), so they are rewritten
// one after the other.  Segments longer than G*WPL words take independent byte pairs.  Every lane of the warp must
// call it (inactive groups pass sl = 0).
template <bool REV, bool LUT, u32 G, u32 WPL>
__device__ __forceinline__ void record_inplace(u8 *d, u32 so, u32 sl, u32 qo, const Lut &lut, u32 gl) {
  u32 *w32 = reinterpret_cast<u32 *>(d);
  const bool full = sl >= 2u * G * WPL && sl + 6u <= 4u * G * WPL;  // both segments hold G*WPL/2 .. G*WPL words
  u32 v[WPL];
  int wi = 0, nv = 0;
#pragma unroll
  for (u32 j = 0; j < WPL; j++) v[j] = 0;
  if (__all_sync(0xffffffffu, full || sl == 0)) {
    if (sl) seg_read_full<REV, LUT, G, WPL>(w32, lut, so, so + sl, gl, v, wi);
    __syncwarp();
    if (sl) seg_write<WPL>(w32, v, wi, (int)WPL);
    if (REV) {
      if (sl) seg_read_full<true, false, G, WPL>(w32, lut, qo, qo + sl, gl, v, wi);
      __syncwarp();
      if (sl) seg_write<WPL>(w32, v, wi, (int)WPL);
    }
  } else if (!__any_sync(0xffffffffu, sl + 6u > G * WPL * 4u)) {
    seg_read<REV, LUT, WPL>(w32, lut, so, so + sl, gl, v, wi, nv);
    __syncwarp();
    seg_write<WPL>(w32, v, wi, nv);
    if (REV) {
      seg_read<true, false, WPL>(w32, lut, qo, qo + sl, gl, v, wi, nv);
      __syncwarp();
      seg_write<WPL>(w32, v, wi, nv);
    }
  } else {
    // long segments: independent byte pairs (i, L-1-i), no hazards
    if (REV) {
      const u32 half = sl >> 1;
      for (u32 i = gl; i < half; i += G) {
        u32 x = d[so + i], y = d[so + sl - 1 - i];
        if (LUT) { x = lut.one(x); y = lut.one(y); }
        d[so + i] = (u8)y;
        d[so + sl - 1 - i] = (u8)x;
        const u8 p = d[qo + i], q = d[qo + sl - 1 - i];
        d[qo + i] = q;
        d[qo + sl - 1 - i] = p;
      }
      if (LUT && (sl & 1u) && gl == 0) d[so + half] = (u8)lut.one(d[so + half]);
    } else if (LUT) {
      for (u32 i = gl; i < sl; i += G) d[so + i] = (u8)lut.one(d[so + i]);
    }
    __syncwarp();
  }
}

// Owned records of a tile, G lanes per record: record r opens at line kmin + 4r (all of them have their four lines in
// the list).  The group checks its record against the grammar of SeqParser.Read -- "@h \n s \n + \n q \n" with
// |s| == |q|, followed by a record line or the end of the file -- writes its element slot and rewrites it in place.
// A record that is anything else raises sm.bad_rec (the block then goes to the general path) and is left alone.
template <u32 NT, u32 G, u32 WPL, class SM, class IsStart>
__device__ __forceinline__ void transform_tile(SM &sm, u8 *d, u32 kmin, u32 n_own, int reverse, int use_lut, u16 *slots,
                                               bool eof, u32 lim, IsStart is_start) {
  const u32 g = threadIdx.x / G, gl = threadIdx.x % G;
  const Lut lut(sm.lut);
  for (u32 rb = 0; rb < n_own; rb += NT / G) {  // uniform trip count per CTA
    const u32 r = rb + g;
    u32 so = 0, sl = 0, qo = 0;
    if (r < n_own) {
      const u32 k = kmin + 4u * r;
      const u32 l0 = sm.ls[k], l1 = sm.ls[k + 1], l2 = sm.ls[k + 2], l3 = sm.ls[k + 3], l4 = sm.ls[k + 4];
      const u32 ql = l4 - 1u - l3;
      so = l1;
      sl = l2 - 1u - l1;
      qo = l3;
      bool ok = d[l1] != '@';                             // the sequence line does not open a record
      ok = ok && (l3 - l2 == 2) && d[l2] == '+';          // bare "+" line
      ok = ok && sl == ql && !(sl > 0 && d[l1] == '+');   // a sequence line starting with '+' flips the parser
      ok = ok && (is_start(l4) || (eof && l4 >= lim));    // next line opens a record, or the file ends here
      if (!ok) {
        sm.bad_rec = 1;  // (not sm.bad: slower warps may still be reading that one behind the counting barrier)
        sl = 0;
      } else if (gl == 0) {
        slots[r] = (u16)l0;  // element slot, input order
      }
    }
    if (reverse) {
      if (use_lut) record_inplace<true, true, G, WPL>(d, so, sl, qo, lut, gl);
      else record_inplace<true, false, G, WPL>(d, so, sl, qo, lut, gl);
    } else if (use_lut) {
      record_inplace<false, true, G, WPL>(d, so, sl, qo, lut, gl);
    }
  }
}

template <class C>
__global__ void __launch_bounds__(C::NT, C::CTAS) k_fastq_inplace(FqInplaceArgs a) {
  using namespace fq;
  constexpr u32 NT = C::NT, NWARP = C::NWARP, T = C::T, LCAP = C::LCAP;
  typedef typename C::Smem Smem;
#ifndef BSK_EMU
  extern __shared__ __align__(256) unsigned char fq_raw_smem[];
  Smem *smp = reinterpret_cast<Smem *>(fq_raw_smem);
#else
  BSK_DYN_SMEM(Smem, smp);
#endif
  Smem &sm = *smp;
  const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const u32 n = a.n, n16 = n & ~15u;

  if (tid < 256) sm.lut[tid] = a.lut[tid];
  if (tid == 0) {
    for (u32 s = 0; s < NSTAGE; s++) tma::mbar_init(&sm.full[s], 1);
    tma::fence_barrier_init();
  }
  __syncthreads();

  // bulk load of the region of tile `tile` into stage s; bytes may be 0 (then there is no barrier phase).
  // n < 4 GiB - 1 MiB (engine.h kMaxBlockBytes), so t0 + T + H does not wrap.
  auto bulk_range = [&](u32 tile, u32 &g0, u32 &g1) {
    const u32 t0 = tile * T;
    g0 = tile ? t0 - PRE : 0u;
    g1 = t0 + T + H < n16 ? t0 + T + H : n16;
    return g1 > g0;
  };
  auto issue = [&](u32 tile, u32 s) {
    u32 g0, g1;
    if (bulk_range(tile, g0, g1)) {
      tma::mbar_expect_tx(&sm.full[s], g1 - g0);
      tma::bulk_load(&sm.in[s][g0 + PRE - tile * T], a.in + g0, g1 - g0, &sm.full[s]);
    }
  };

  if (tid == 0) {
    for (u32 p = 0; p + 1 < NSTAGE; p++) {
      const u32 tl = blockIdx.x + p * gridDim.x;
      if (tl < a.n_tiles) issue(tl, p);
    }
  }

  u32 it = 0;
  for (u32 tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, it++) {
    const u32 s = it % NSTAGE;
    const u32 parity = (it / NSTAGE) & 1u;
    const u32 t0 = tile * T;
    u8 *stage = sm.in[s];
    u8 *d = stage + PRE;  // region byte 0 == global byte t0
    const u32 lim = (n - t0 < T + H) ? n - t0 : T + H;  // valid bytes of the region
    const bool eof = (n - t0) <= T + H;
    {
      u32 g0, g1;
      if (bulk_range(tile, g0, g1)) tma::mbar_wait(&sm.full[s], parity);
    }
    // bytes the bulk copy did not bring: the look-behind of tile 0, the ragged tail of the file, padding
    if (tile == 0 || t0 + T + H > n16) {
      for (u32 i = tid; i < PRE + T + H + 16; i += NT) {
        const bool before = t0 + i < PRE;  // only tile 0
        const u32 g = t0 + i - PRE;
        if (before || g >= n16) stage[i] = (!before && g < n) ? a.in[g] : (u8)'\n';
      }
      __syncthreads();
    }

    // record-start rule: a line opens a record iff it starts with '@' unless the line before it is a
    // bare "+" ("\n+\n@" is the quality line of a record)
    auto is_start = [&](u32 p) {
      return p < lim && d[p] == '@' && !(t0 + p >= 3u && d[(int)p - 3] == '\n' && d[(int)p - 2] == '+');
    };

    // ---- newline scan -> line starts, first record line, grammar check of the owned records.
    // The halo is scanned only as far as records usually reach (a.scan_halo); a tile whose last owned record
    // ends beyond that repeats the scan over the whole halo.
    u32 hs = a.scan_halo;
    u32 n_lines = 0, n_own = 0, kmin = 0;
    bool bad = false;
    for (;;) {  // uniform
      const u32 slim = lim < T + hs ? lim : T + hs;  // bytes scanned
      const u32 span = tid * (CPL * 16u);
      u32 mlo = 0, mhi = 0;
      if (warp * (32u * CPL * 16u) < slim) {
        // lane owns CPL consecutive 16-byte chunks; the 0x80 flag bytes of a chunk are packed into a position-
        // ordered 16-bit mask with four IDP.4A (flag byte * {1,2,4,8} summed = nibble << 7)
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
      if (lane == 31) sm.wtot[warp] = inc;
      if (tid == 0) { sm.bad = 0; sm.rescan = 0; sm.kmin = 0xffffffffu; sm.bad_rec = 0; }
      __syncthreads();
      u32 base, n_nl;
      {
        constexpr u32 W = NWARP <= 16 ? 16 : 32;  // lanes that scan the warp totals
        u32 x = (lane & (W - 1u)) < NWARP ? sm.wtot[lane & (W - 1u)] : 0u;
#pragma unroll
        for (int off = 1; off < (int)W; off <<= 1) {
          const u32 y = __shfl_up_sync(0xffffffffu, x, off, W);
          if ((int)(lane & (W - 1u)) >= off) x += y;
        }
        n_nl = __shfl_sync(0xffffffffu, x, W - 1);
        base = __shfl_sync(0xffffffffu, x, (warp + W - 1u) & (W - 1u));
        if (warp == 0) base = 0;
      }
      const bool virt = eof && slim == lim && lim > 0 && d[lim - 1] != '\n';  // unterminated last line
      if (n_nl + 2 > LCAP) { bad = true; break; }
      {
        u32 k = base + inc - cnt + 1;  // ls[k] = start of the line after the (k-1)-th newline
        while (mlo | mhi) {
          u32 t;
          if (mlo) { t = (u32)__ffs((int)mlo) - 1u; mlo &= mlo - 1u; }
          else { t = 32u + (u32)__ffs((int)mhi) - 1u; mhi &= mhi - 1u; }
          const u32 p = span + t + 1u;
          if (k < 8u && is_start(p)) atomicMin(&sm.kmin, k);  // first record line: one of lines 0..4 in a 4-line stream
          sm.ls[k++] = (u16)p;
        }
        if (tid == 0) {
          sm.ls[0] = 0;
          if (virt) sm.ls[n_nl + 1] = (u16)(lim + 1);
          if (lim > 0 && d[0] == '@' && (tile == 0 || (d[-1] == '\n' && !(d[-3] == '\n' && d[-2] == '+')))) sm.kmin = 0;
        }
      }
      n_lines = n_nl + (virt ? 1u : 0u);  // ls[0 .. n_lines] are valid
      __syncthreads();

      // Owned records = record lines that start inside the tile.  In a 4-line stream record r opens at line
      // kmin + 4r (kmin: the first record line, found while the list was written); the chain is checked record by
      // record in the transform pass below.  Here only: how many are there, and does the last one end inside the
      // scanned part of the halo?
      kmin = sm.kmin;
      bool own = false;
      if (kmin != 0xffffffffu) {
        const u32 k = kmin + 4u * tid;
        if (k <= n_lines) {
          const u32 l0 = sm.ls[k];
          if (l0 < T && l0 < lim) {  // (the entry behind the last line of the file is not a line)
            own = true;
            if (k + 4 > n_lines) {
              if (slim < lim) sm.rescan = 1;  // the record ends beyond the scanned part of the halo
              else sm.bad = 1;                // longer than the halo, or truncated
            }
          }
        }
      } else if (n_lines >= 8u && tid == 0) {
        sm.bad = 1;  // eight lines without a record line: not a 4-line stream
      }
      n_own = (u32)__syncthreads_count(own);
      bad = sm.bad != 0 || n_own >= NT;  // (a thread per chain position: NT of them may not be all)
      if (tile == 0 && kmin != 0) bad = true;  // the file must open with a marked record
      const bool rescan = sm.rescan != 0;
      if (bad || !rescan || hs >= H) {
        if (rescan) bad = true;
        break;
      }
      hs = H;
      __syncthreads();  // everybody has read the flags before thread 0 clears them again
    }
    // refill the stage the PREVIOUS tile was stored from (its bulk store has long finished reading by now)
    if (tid == 0) {
      const u32 tn = tile + (NSTAGE - 1) * gridDim.x;
      if (it > 0) tma::bulk_wait_read();
      if (tn < a.n_tiles) issue(tn, (it + NSTAGE - 1) % NSTAGE);
    }
    if (bad) {
      if (tid == 0) {
        atomicAdd((unsigned long long *)&a.st->counters[0], 1ull);
        // diagnostics: first tile outside the grammar and what the CTA saw there
        const unsigned long long info = ((unsigned long long)tile << 32) | ((u64)(sm.bad & 1u) << 31) | ((u64)(sm.rescan & 1u) << 30) |
                                        ((u64)(n_own & 0x3ffu) << 16) | (n_lines & 0xffffu);
        atomicMin((unsigned long long *)&a.st->counters[4], info);
      }
      __syncthreads();
      continue;  // uniform
    }

    // ---- grammar check + in-place transform, record by record
    if (n_own) {
      u16 *slots = a.slots + (size_t)tile * RCAP;
      if (a.group == 4) {
        transform_tile<NT, 4, 10>(sm, d, kmin, n_own, a.reverse, a.use_lut, slots, eof, lim, is_start);
      } else if (a.group == 8) {
        if (a.wpl <= 5) transform_tile<NT, 8, 5>(sm, d, kmin, n_own, a.reverse, a.use_lut, slots, eof, lim, is_start);
        else transform_tile<NT, 8, 8>(sm, d, kmin, n_own, a.reverse, a.use_lut, slots, eof, lim, is_start);
      } else {
        transform_tile<NT, 32, 4>(sm, d, kmin, n_own, a.reverse, a.use_lut, slots, eof, lim, is_start);
      }
    }
    // ---- owned byte range [lo, hi) -> out, same offsets: bulk store of the aligned body, ragged ends by warp 0.
    // Owned records are contiguous: from the first owned record line to the line after the last one.
    // Only warp 0 waits for the transform to finish; the other warps arrive and go on to scan the next tile (other
    // stage; nothing of this tile is overwritten before the next full barrier, which warp 0 joins after the store).
    tma::fence_proxy_async();
    if (warp != 0) {
      tma::named_arrive(1, NT);
    } else {
      tma::named_sync(1, NT);
    }
    if (warp == 0 && sm.bad_rec) {  // a record outside the grammar: the block goes to the general path
      if (lane == 0) {
        atomicAdd((unsigned long long *)&a.st->counters[0], 1ull);
        const unsigned long long info = ((unsigned long long)tile << 32) | (1ull << 31) | ((u64)(n_own & 0x3ffu) << 16) | (n_lines & 0xffffu);
        atomicMin((unsigned long long *)&a.st->counters[4], info);
      }
    } else if (warp == 0 && n_own > 0) {
      const u32 lo = sm.ls[kmin];
      const u32 hi = sm.ls[kmin + 4u * n_own];  // start of the next record == one past the '\n' that ends the last owned one
      const u32 lo16 = (lo + 15u) & ~15u, hi16 = hi & ~15u;
      u8 *go = a.out + t0;
      if (hi16 > lo16) {
        for (u32 i = lo + lane; i < lo16; i += 32) go[i] = d[i];
        for (u32 i = hi16 + lane; i < hi; i += 32) go[i] = d[i];
        __syncwarp();
        if (lane == 0) {
          tma::bulk_store(go + lo16, d + lo16, hi16 - lo16);
          tma::bulk_commit();
        }
      } else {
        for (u32 i = lo + lane; i < hi; i += 32) go[i] = d[i];
        __syncwarp();
      }
      if (eof && lane == 0 && hi >= lim) a.st->counters[1] = (u64)t0 + hi;  // total output bytes (n, or n + 1)
    }
    if (tid == 0) a.tile_cnt[tile] = n_own;
    // no barrier here: the next tile's first barrier comes before anything above is overwritten
  }
  if (tid == 0) tma::bulk_wait_all();
}

// element offsets from the per-tile slot lists: one warp per tile.  Launched before the host knows the record
// count: writes are bounded by `cap` entries (the host re-runs it with a larger array in the rare overflow case),
// and the terminating entry elem_off[n_records] = total output bytes comes from the main kernel's status word.
__global__ void k_fastq_elem_expand(const u32 *__restrict__ tile_cnt, const u64 *__restrict__ tile_base,
                                    const u16 *__restrict__ slots, u64 *__restrict__ elem_off, u32 n_tiles, u32 tile_bytes,
                                    u64 cap, const DevStatus *__restrict__ st) {
  const u32 tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (tile >= n_tiles) return;
  const u32 c = tile_cnt[tile];
  const u64 b = tile_base[tile];
  for (u32 r = lane; r < c; r += 32)
    if (b + r < cap) elem_off[b + r] = (u64)tile * tile_bytes + slots[(size_t)tile * fq::RCAP + r];
  if (tile == n_tiles - 1 && lane == 0) {
    const u64 nrec = tile_base[n_tiles];
    if (nrec < cap) elem_off[nrec] = st->counters[1];
  }
}

// CTA shape of the process: 512 x 4 unless BSK_FQ_SHAPE says otherwise (A/B runs: profiles/r2_experiments.txt)
static int fq_shape() {
  const char *e = getenv("BSK_FQ_SHAPE");
  const int v = e ? atoi(e) : 0;
  return v >= 0 && v <= 3 ? v : 0;
}
u32 fastq_inplace_tile_bytes() {
  switch (fq_shape()) {
    case 1: return fq::CfgB::T;
    case 2: return fq::CfgC::T;
    case 3: return fq::CfgD::T;
    case 4: return fq::CfgE::T;
    default: return fq::CfgA::T;
  }
}
u32 fastq_inplace_tiles(u32 n) { return (n + fastq_inplace_tile_bytes() - 1) / fastq_inplace_tile_bytes(); }
u32 fastq_inplace_slot_stride() { return fq::RCAP; }

template <class C>
static void launch_fq(const FqInplaceArgs &a, int n_sm, cudaStream_t s) {
  const size_t smem = sizeof(typename C::Smem) + 16;
#ifndef BSK_EMU
  // the opt-in to > 48 KiB of dynamic shared memory is per device (a process may hold ctxs on several GPUs)
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaFuncSetAttribute(k_fastq_inplace<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set[dev] = true;
  }
#endif
  u32 grid = (u32)n_sm * C::CTAS;
  if (grid > a.n_tiles) grid = a.n_tiles;
  if (grid == 0) return;
  BSK_LAUNCH(k_fastq_inplace<C>, grid, C::NT, smem, s, a);
}

void fastq_inplace(const u8 *in, u32 n, u8 *out, const u8 *lut, u32 *tile_cnt, u16 *slots, DevStatus *st, int reverse,
                   int use_lut, int group, u32 max_seg, u32 scan_halo, int n_sm, cudaStream_t s) {
  FqInplaceArgs a;
  a.in = in;
  a.n = n;
  a.out = out;
  a.lut = lut;
  a.tile_cnt = tile_cnt;
  a.slots = slots;
  a.st = st;
  a.n_tiles = fastq_inplace_tiles(n);
  a.reverse = reverse;
  a.use_lut = use_lut;
  a.group = group == 8 ? 8 : (group == 4 ? 4 : 32);
  a.wpl = max_seg <= 154 ? 5 : 8;
  scan_halo = (scan_halo + 15u) & ~15u;
  a.scan_halo = scan_halo < 256u ? 256u : (scan_halo > fq::H ? fq::H : scan_halo);
  switch (fq_shape()) {
    case 1: launch_fq<fq::CfgB>(a, n_sm, s); break;
    case 2: launch_fq<fq::CfgC>(a, n_sm, s); break;
    case 3: launch_fq<fq::CfgD>(a, n_sm, s); break;
    case 4: launch_fq<fq::CfgE>(a, n_sm, s); break;
    default: launch_fq<fq::CfgA>(a, n_sm, s);
  }
}

void fastq_elem_expand(const u32 *tile_cnt, const u64 *tile_base, const u16 *slots, u64 *elem_off, u32 n_tiles, u64 cap,
                       const DevStatus *st, cudaStream_t s) {
  if (!n_tiles) return;
  const u64 threads = (u64)n_tiles * 32;
  BSK_LAUNCH_FLAT(k_fastq_elem_expand, (u32)((threads + 255) / 256), 256, 0, s, tile_cnt, tile_base, slots, elem_off, n_tiles,
                  fastq_inplace_tile_bytes(), cap, st);
}

}  // namespace k
}  // namespace bsk
