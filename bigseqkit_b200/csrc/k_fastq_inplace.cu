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
//   * persistent CTAs walk 20 KiB tiles (+ 4 KiB halo) with a 3-stage ring filled by 1-D TMA bulk loads;
//   * a CTA finds the newlines of its region (16-byte shared-memory loads, SWAR zero-byte test), turns
//     them into line starts, classifies record starts, checks every record it owns (the ones that START
//     inside the tile) against the 4-line grammar of SeqParser.Read;
//   * the sequence (reverse + 256-entry byte map) and the quality (reverse) of every owned record are
//     rewritten IN PLACE in the stage buffer, 4 bytes per lane with PRMT, 16 lanes per record;
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
constexpr u32 RCAP = 512;      // owned records per tile (slot stride)

// CTA shape: NT threads, every lane scans CPL consecutive 16-byte chunks, so tile + halo = NT * CPL * 16 bytes
template <u32 NT_, u32 CPL_, u32 CTAS_, u32 NSTAGE_, u32 LCAP_ = 3072>
struct Cfg {
  static constexpr u32 NT = NT_, CPL = CPL_, CTAS = CTAS_, NSTAGE = NSTAGE_;
  static constexpr u32 LCAP = LCAP_;                          // line starts per region
  static constexpr u32 NWARP = NT / 32;
  static constexpr u32 T = NT * CPL * 16 - H;                 // tile bytes
  static constexpr u32 STAGE = PRE + T + H + 16;
  static constexpr u32 LITER = (LCAP + NT) / NT;              // passes of the CTA over the line list
  static_assert(T % 16 == 0 && STAGE % 16 == 0 && T + H < 32768, "tile geometry");
  struct Smem {
    u8 in[NSTAGE][STAGE];
    u64 full[NSTAGE];
    u8 lut[256];
    u16 ls[LCAP + 8];          // line starts (bit 15: the line opens a record), ls[0] = 0
    u16 r_line[RCAP];          // first line of every owned record, any order
    u32 wtot[NWARP];
    u32 wtot2[LITER][NWARP];   // owned records per (pass over the line list, warp)
    u32 bad, rescan, n_list, kmin, kmax;
  };
};
typedef Cfg<512, 3, 2, 3> CfgA;   // 20 KiB tiles, 2 CTAs / SM, 3-stage ring
typedef Cfg<256, 4, 4, 3> CfgB;   // 12 KiB tiles, 4 CTAs / SM: smaller barrier domains (3 resident: shared memory)
typedef Cfg<512, 3, 2, 4> CfgC;   // as A with a 4-stage ring (loads issued 2.6 tiles ahead)
typedef Cfg<512, 3, 3, 2> CfgD;   // as A with a 2-stage ring and 3 CTAs / SM (<= 40 registers per thread)
typedef Cfg<512, 3, 4, 2, 2048> CfgE;   // 4 CTAs / SM: 64 warps, <= 32 registers per thread, 55 KB of shared memory
typedef Cfg<384, 4, 4, 2, 2048> CfgF;   // 4 CTAs / SM of 12 warps, <= 40 registers per thread
typedef Cfg<448, 3, 3, 3> CfgG;         // 17 KiB tiles: 3 CTAs / SM of 14 warps with a 3-stage ring, <= 48 registers
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
  int group;       // lanes per record in the transform: 8 (records <= ~250 B per segment), 16, 32
  int wpl;         // group == 8: words per lane (5 covers segments up to 160 B, else 8)
  int issue_late;  // refill the ring after the grammar check (default) instead of at the top of the tile (BSK_FQ_EARLY)
  u32 scan_halo;   // bytes of the halo the newline scan covers on its first attempt (multiple of 16, <= H)
};

// flags (0x80 per byte) of the bytes of w that equal '\n'; exact for every byte value
__device__ __forceinline__ u32 nl_flags(u32 w) {
  const u32 x = w ^ 0x0a0a0a0au;
  const u32 y = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
  return ~(y | x) & 0x80808080u;
}

__device__ __forceinline__ u32 lut4(const u8 *lut, u32 v) {
  return (u32)lut[v & 0xffu] | ((u32)lut[(v >> 8) & 0xffu] << 8) | ((u32)lut[(v >> 16) & 0xffu] << 16) |
         ((u32)lut[v >> 24] << 24);
}

// Word geometry of the in-place rewrite of the byte range [a, e): whole 32-bit words [ai, ae) are produced with
// one PRMT from two source words, the <= 3 ragged bytes at each end go through a byte path.
struct SegGeo {
  u32 a, e, nwf, sel, q0, wi;
  __device__ __forceinline__ SegGeo(u32 a_, u32 L) {
    a = a_;
    e = a_ + L;
    const u32 ai = (a + 3u) & ~3u, ae = e & ~3u;
    nwf = ae > ai ? (ae - ai) >> 2 : 0u;
    const u32 U0 = a + e - 4u - ai;  // lowest source byte of interior word 0 (dest byte ai+b <- source U0+3-b)
    const u32 sh = U0 & 3u;
    sel = 0x0123u + sh * 0x1111u;    // PRMT selector {sh+3, sh+2, sh+1, sh}
    q0 = U0 >> 2;
    wi = ai >> 2;
  }
  // ragged byte handled by lane t of the group (t < 3: head, 3 <= t < 6: tail); false when there is none
  __device__ __forceinline__ bool edge(u32 t, u32 &x) const {
    const u32 ai = wi << 2, ae = e & ~3u;
    const u32 head_end = ai < e ? ai : e;
    const u32 tail_beg = ae > ai ? ae : head_end;
    x = t < 3u ? a + t : tail_beg + (t - 3u);
    return t < 3u ? x < head_end : (t < 6u && x < e);
  }
};

// In-place rewrite of one record by a group of G lanes: sequence [so, so+sl) reversed (REV) and mapped through
// lut (LUT), quality [qo, qo+sl) reversed (REV).  Registers hold both segments until every lane has read, so the
// rewrite is hazard-free; segments longer than G*WPL words take independent byte pairs.  Every lane of the warp
// must call it (inactive groups pass sl = 0).
template <bool REV, bool LUT, u32 G, u32 WPL>
__device__ __forceinline__ void record_inplace(u8 *d, u32 so, u32 sl, u32 qo, const u8 *lut, u32 gl) {
  u32 *w32 = reinterpret_cast<u32 *>(d);
  const SegGeo S(so, sl), Q(qo, sl);
  const bool slow = __any_sync(0xffffffffu, S.nwf > G * WPL || Q.nwf > G * WPL);
  if (!slow) {
    constexpr u32 ER = (6 + G - 1) / G;  // rounds a group needs to cover the six ragged bytes of a segment
    u32 xs[ER], xq[ER];
    bool es[ER], eq[ER];
    u8 vs[ER], vq[ER];
#pragma unroll
    for (u32 t = 0; t < ER; t++) {
      es[t] = S.edge(gl + t * G, xs[t]);
      eq[t] = REV && Q.edge(gl + t * G, xq[t]);
      vs[t] = 0;
      vq[t] = 0;
      if (es[t]) {
        vs[t] = d[REV ? (S.a + S.e - 1u - xs[t]) : xs[t]];
        if (LUT) vs[t] = lut[vs[t]];
      }
      if (eq[t]) vq[t] = d[Q.a + Q.e - 1u - xq[t]];
    }
    u32 sv[WPL], qv[WPL];
#pragma unroll
    for (u32 j = 0; j < WPL; j++) {
      const u32 idx = gl + j * G;
      sv[j] = 0;
      qv[j] = 0;
      if (idx < S.nwf) {
        if (REV) sv[j] = __byte_perm(w32[S.q0 - idx], w32[S.q0 - idx + 1u], S.sel);
        else sv[j] = w32[S.wi + idx];
      }
      if (REV && idx < Q.nwf) qv[j] = __byte_perm(w32[Q.q0 - idx], w32[Q.q0 - idx + 1u], Q.sel);
      if (LUT && idx < S.nwf) sv[j] = lut4(lut, sv[j]);
    }
    __syncwarp();
#pragma unroll
    for (u32 j = 0; j < WPL; j++) {
      const u32 idx = gl + j * G;
      if (idx < S.nwf) w32[S.wi + idx] = sv[j];
      if (REV && idx < Q.nwf) w32[Q.wi + idx] = qv[j];
    }
#pragma unroll
    for (u32 t = 0; t < ER; t++) {
      if (es[t]) d[xs[t]] = vs[t];
      if (eq[t]) d[xq[t]] = vq[t];
    }
    __syncwarp();
  } else {
    // long segments: independent byte pairs (i, L-1-i), no hazards
    if (REV) {
      const u32 half = sl >> 1;
      for (u32 i = gl; i < half; i += G) {
        u8 x = d[so + i], y = d[so + sl - 1 - i];
        if (LUT) { x = lut[x]; y = lut[y]; }
        d[so + i] = y;
        d[so + sl - 1 - i] = x;
        const u8 p = d[qo + i], q = d[qo + sl - 1 - i];
        d[qo + i] = q;
        d[qo + sl - 1 - i] = p;
      }
      if (LUT && (sl & 1u) && gl == 0) d[so + half] = lut[d[so + half]];
    } else if (LUT) {
      for (u32 i = gl; i < sl; i += G) d[so + i] = lut[d[so + i]];
    }
    __syncwarp();
  }
}

// transform of all owned records of a tile, G lanes per record (r_line holds them in any order)
template <class C, u32 G, u32 WPL>
__device__ __forceinline__ void transform_tile(typename C::Smem &sm, u8 *d, u32 n_own, int reverse, int use_lut) {
  const u32 g = threadIdx.x / G, gl = threadIdx.x % G;
  for (u32 rb = 0; rb < n_own; rb += C::NT / G) {  // uniform trip count per CTA
    const u32 r = rb + g;
    u32 so = 0, sl = 0, qo = 0;
    if (r < n_own) {
      const u32 k = sm.r_line[r];
      so = sm.ls[k + 1] & 0x7fffu;
      sl = (sm.ls[k + 2] & 0x7fffu) - 1u - so;
      qo = sm.ls[k + 3] & 0x7fffu;
    }
    if (reverse) {
      if (use_lut) record_inplace<true, true, G, WPL>(d, so, sl, qo, sm.lut, gl);
      else record_inplace<true, false, G, WPL>(d, so, sl, qo, sm.lut, gl);
    } else {
      record_inplace<false, true, G, WPL>(d, so, sl, qo, sm.lut, gl);
    }
  }
}

template <class C>
__global__ void __launch_bounds__(C::NT, C::CTAS) k_fastq_inplace(FqInplaceArgs a) {
  using namespace fq;
  constexpr u32 NT = C::NT, CPL = C::CPL, NWARP = C::NWARP, T = C::T, LITER = C::LITER, NSTAGE = C::NSTAGE, LCAP = C::LCAP;
  typedef typename C::Smem Smem;
  BSK_DYN_SMEM(Smem, smp);
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
    if (tid == 0 && !a.issue_late) {
      const u32 tn = tile + (NSTAGE - 1) * gridDim.x;
      if (it > 0) tma::bulk_wait_read();  // the stage being refilled was the source of the previous tile's store
      if (tn < a.n_tiles) issue(tn, (it + NSTAGE - 1) % NSTAGE);
    }
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

    // ---- newline scan -> line starts (bit 15 = the line opens a record), grammar check, owned-record list.
    // The halo is scanned only as far as records usually reach (a.scan_halo); a tile whose last owned record
    // ends beyond that repeats the scan over the whole halo.
    u32 hs = a.scan_halo;
    u32 n_lines = 0, n_own = 0;
    bool bad = false;
    bool own0 = false;            // first pass over the line list: this thread's line opens an owned record
    u32 bal0 = 0, own_pos0 = 0;   //   ... the warp's ballot of that, and the record's start
    for (;;) {  // uniform
      const u32 slim = lim < T + hs ? lim : T + hs;  // bytes scanned
      const u32 span = (warp * 32u + lane) * (CPL * 16u);
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
        static_assert(CPL == 3 || CPL == 4, "mask packing");
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
      if (lane == 31) sm.wtot[warp] = inc;
      if (tid == 0) { sm.bad = 0; sm.rescan = 0; sm.n_list = 0; sm.kmin = 0xffffffffu; sm.kmax = 0; }
      __syncthreads();
      u32 base, n_nl;
      {
        static_assert(NWARP <= 16, "warp totals are scanned by 16 lanes");
        u32 x = (lane & 15u) < NWARP ? sm.wtot[lane & 15u] : 0u;
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
      if (n_nl + 2 > LCAP) { bad = true; break; }
      {
        // record-start rule: a line opens a record iff it starts with '@' unless the line before it
        // is a bare "+" ("\n+\n@" is the quality line of a record)
        u32 k = base + inc - cnt + 1;  // ls[k] = start of the line after the (k-1)-th newline
        while (mlo | mhi) {
          u32 t;
          if (mlo) { t = (u32)__ffs((int)mlo) - 1u; mlo &= mlo - 1u; }
          else { t = 32u + (u32)__ffs((int)mhi) - 1u; mhi &= mhi - 1u; }
          const u32 p = span + t + 1u;
          u32 rs = 0;
          if (p < lim && d[p] == '@') rs = (t0 + p >= 3u && d[(int)p - 3] == '\n' && d[(int)p - 2] == '+') ? 0u : 0x8000u;
          sm.ls[k++] = (u16)(p | rs);
        }
        if (tid == 0) {
          u32 rs0 = 0;
          if (lim > 0 && d[0] == '@') {
            if (tile == 0) rs0 = 0x8000u;
            else if (d[-1] == '\n' && !(d[-3] == '\n' && d[-2] == '+')) rs0 = 0x8000u;
          }
          sm.ls[0] = (u16)rs0;
          if (virt) sm.ls[n_nl + 1] = (u16)(lim + 1);
        }
      }
      n_lines = n_nl + (virt ? 1u : 0u);  // ls[0 .. n_lines] are valid
      __syncthreads();

      // owned records = record lines that start inside the tile; each must be "@h \n s \n + \n q \n" with |s| == |q|
      for (u32 kb = 0; kb <= n_lines; kb += NT) {  // uniform trip count
        const u32 k = kb + tid;
        bool own = false;
        if (kb + warp * 32u > n_lines) {  // this warp's lines are past the end of the list (warp-uniform)
          if (kb == 0) { own0 = false; bal0 = 0; }
          if (lane == 0 && kb / NT < LITER) sm.wtot2[kb / NT][warp] = 0;
          continue;
        }
        if (k <= n_lines) {
          const u32 e0 = sm.ls[k];
          if ((e0 & 0x8000u) && (e0 & 0x7fffu) < T) {
            own = true;
            if (kb == 0) own_pos0 = e0 & 0x7fffu;
            if (k + 4 <= n_lines) {
              const u32 e1 = sm.ls[k + 1], e2 = sm.ls[k + 2], e3 = sm.ls[k + 3], e4 = sm.ls[k + 4];
              const u32 l1 = e1 & 0x7fffu, l2 = e2 & 0x7fffu, l3 = e3 & 0x7fffu, l4 = e4 & 0x7fffu;
              const u32 sl = l2 - 1 - l1, ql = l4 - 1 - l3;
              bool ok = ((e1 | e2 | e3) & 0x8000u) == 0;
              ok = ok && (l3 - l2 == 2) && d[l2] == '+';         // bare "+" line
              ok = ok && sl == ql && !(sl > 0 && d[l1] == '+');  // a sequence line starting with '+' flips the parser
              ok = ok && ((e4 & 0x8000u) || (eof && l4 >= lim)); // next line opens a record, or the file ends here
              if (!ok) sm.bad = 1;
            } else if (slim < lim) {
              sm.rescan = 1;  // the record ends beyond the scanned part of the halo
            } else {
              sm.bad = 1;     // longer than the halo, or truncated
            }
          }
        }
        const u32 bal = __ballot_sync(0xffffffffu, own);
        if (kb == 0) { own0 = own; bal0 = bal; }
        u32 wpos = 0;
        if (lane == 0) {
          const u32 c = (u32)__popc(bal);
          if (kb / NT < LITER) sm.wtot2[kb / NT][warp] = c;
          if (c) {
            wpos = atomicAdd(&sm.n_list, c);
            atomicMin(&sm.kmin, kb + warp * 32u + (u32)__ffs((int)bal) - 1u);   // lines of the first / last owned record
            atomicMax(&sm.kmax, kb + warp * 32u + 31u - (u32)__clz((int)bal));
          }
        }
        wpos = __shfl_sync(0xffffffffu, wpos, 0);
        if (own) {
          const u32 r = wpos + (u32)__popc(bal & ((1u << lane) - 1u));
          if (r < RCAP) sm.r_line[r] = (u16)k;
        }
      }
      __syncthreads();
      n_own = sm.n_list;
      bad = sm.bad != 0 || n_own > RCAP;
      if (tile == 0 && !(sm.ls[0] & 0x8000u)) bad = true;  // the file must open with a marked record
      const bool rescan = sm.rescan != 0;
      if (bad || !rescan || hs >= H) {
        if (rescan) bad = true;
        break;
      }
      hs = H;
      __syncthreads();  // everybody has read the flags before thread 0 clears them again
    }
    // refill the stage the PREVIOUS tile was stored from (its bulk store has long finished reading by now)
    if (tid == 0 && a.issue_late) {
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

    // ---- in-place transform
    if (a.reverse || a.use_lut) {
      if (a.group == 4) {
        transform_tile<C, 4, 10>(sm, d, n_own, a.reverse, a.use_lut);
      } else if (a.group == 8) {
        if (a.wpl <= 5) transform_tile<C, 8, 5>(sm, d, n_own, a.reverse, a.use_lut);
        else transform_tile<C, 8, 8>(sm, d, n_own, a.reverse, a.use_lut);
      } else if (a.group == 16) {
        transform_tile<C, 16, 4>(sm, d, n_own, a.reverse, a.use_lut);
      } else {
        transform_tile<C, 32, 4>(sm, d, n_own, a.reverse, a.use_lut);
      }
    }
    tma::fence_proxy_async();
    __syncthreads();

    // ---- owned byte range [lo, hi) -> out, same offsets: bulk store of the aligned body, ragged ends by warp 0.
    // Owned records are contiguous: from the first owned record line to the line after the last one.
    if (warp == 0 && n_own > 0) {
      const u32 kmin = sm.kmin, kmax = sm.kmax;
      const u32 lo = sm.ls[kmin] & 0x7fffu;
      const u32 hi = sm.ls[kmax + 4] & 0x7fffu;  // start of the next record == one past the '\n' that ends the last owned one
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

    // ---- element slots in input order: rank of a record = owned records on earlier lines
    if (n_lines < NT) {  // one pass over the line list (the usual case): ownership is still in registers
      if (bal0) {
        u32 wb = 0;
        for (u32 w = 0; w < warp; w++) wb += sm.wtot2[0][w];
        if (own0) a.slots[(size_t)tile * RCAP + wb + (u32)__popc(bal0 & ((1u << lane) - 1u))] = (u16)own_pos0;
      }
    } else {
      u32 before = 0;  // owned records of earlier passes over the line list
      for (u32 kb = 0; kb <= n_lines; kb += NT) {  // uniform trip count
        const u32 it2 = kb / NT;
        const u32 k = kb + tid;
        bool own = false;
        u32 e0 = 0;
        if (k <= n_lines) {
          e0 = sm.ls[k];
          own = (e0 & 0x8000u) && (e0 & 0x7fffu) < T;
        }
        const u32 bal = __ballot_sync(0xffffffffu, own);
        u32 wb = 0, tot = 0;
        if (it2 < LITER) {
          for (u32 w = 0; w < NWARP; w++) {
            const u32 c = sm.wtot2[it2][w];
            if (w < warp) wb += c;
            tot += c;
          }
        }
        if (own) a.slots[(size_t)tile * RCAP + before + wb + (u32)__popc(bal & ((1u << lane) - 1u))] = (u16)(e0 & 0x7fffu);
        before += tot;
      }
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

u32 fastq_inplace_tile_bytes(int variant) {  // all others share A's tile
  return variant == 1 ? fq::CfgB::T : variant == 6 ? fq::CfgG::T : fq::CfgA::T;
}
static_assert(fq::CfgC::T == fq::CfgA::T && fq::CfgD::T == fq::CfgA::T && fq::CfgE::T == fq::CfgA::T && fq::CfgF::T == fq::CfgA::T, "tile bytes");
u32 fastq_inplace_tiles(u32 n, int variant) {
  const u32 t = fastq_inplace_tile_bytes(variant);
  return (n + t - 1) / t;
}
u32 fastq_inplace_slot_stride() { return fq::RCAP; }

template <class C>
static void launch_fastq_inplace(FqInplaceArgs a, int n_sm, cudaStream_t s) {
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
                   int use_lut, int group, u32 max_seg, u32 scan_halo, int variant, int n_sm, cudaStream_t s) {
  FqInplaceArgs a;
  a.in = in;
  a.n = n;
  a.out = out;
  a.lut = lut;
  a.tile_cnt = tile_cnt;
  a.slots = slots;
  a.st = st;
  a.n_tiles = fastq_inplace_tiles(n, variant);
  a.reverse = reverse;
  a.use_lut = use_lut;
  a.group = group;
  a.issue_late = getenv("BSK_FQ_EARLY") ? 0 : 1;
  a.wpl = max_seg <= 157 ? 5 : 8;
  scan_halo = (scan_halo + 15u) & ~15u;
  a.scan_halo = scan_halo < 256u ? 256u : (scan_halo > fq::H ? fq::H : scan_halo);
  if (variant == 1) launch_fastq_inplace<fq::CfgB>(a, n_sm, s);
  else if (variant == 2) launch_fastq_inplace<fq::CfgC>(a, n_sm, s);
  else if (variant == 3) launch_fastq_inplace<fq::CfgD>(a, n_sm, s);
  else if (variant == 4) launch_fastq_inplace<fq::CfgE>(a, n_sm, s);
  else if (variant == 5) launch_fastq_inplace<fq::CfgF>(a, n_sm, s);
  else if (variant == 6) launch_fastq_inplace<fq::CfgG>(a, n_sm, s);
  else launch_fastq_inplace<fq::CfgA>(a, n_sm, s);
}

void fastq_elem_expand(const u32 *tile_cnt, const u64 *tile_base, const u16 *slots, u64 *elem_off, u32 n_tiles, int variant,
                       u64 cap, const DevStatus *st, cudaStream_t s) {
  if (!n_tiles) return;
  const u64 threads = (u64)n_tiles * 32;
  BSK_LAUNCH_FLAT(k_fastq_elem_expand, (u32)((threads + 255) / 256), 256, 0, s, tile_cnt, tile_base, slots, elem_off, n_tiles,
                  fastq_inplace_tile_bytes(variant), cap, st);
}

}  // namespace k
}  // namespace bsk
