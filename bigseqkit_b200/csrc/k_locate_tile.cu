// k_locate_tile.cu -- `locate` with a panel of equal-length ACGT patterns on FASTA, in ONE streaming pass over the
// raw input (BASELINE configs[3]: 1000 x 12-mer on contigs wrapped at 60).
//
//   Locate.Call, default exact path   bigseqkit-lib/locate.go:395-769: for every pattern (and, unless -P, on the
//     reverse complement too) all occurrences by repeated bytes.Index, greedy step start+1 (:583-667, :669-766)
//   SeqParser.Read (FASTA)            bigseqkit-lib/helper.go:240-250: the sequence is every line after the header,
//                                     joined without their '\n'
//
// The reference makes (#patterns x 2) passes over every sequence and builds a fresh reverse complement per pattern.
// Here the input is read once, newlines in place:
//   * persistent CTAs (512 threads, 4 per SM) walk 23 KiB tiles with a 1 KiB look-behind, staged by 1-D TMA bulk loads;
//   * a SWAR newline scan gives every lane its newline masks (-> sequence coordinates) and finds the header lines
//     (a '>' that follows a newline), whose bytes are overwritten in the stage buffer so that no window spans them;
//   * every lane packs its 48 bytes (16 bytes of warm-up) into 2-bit codes, a 16-byte chunk at a time (SWAR: code =
//     bits 1-2 of the letter, validity by rebuilding the letter from the code; a single newline is cut out of the
//     packed word), keeps the codes of the last 16 bases beside and takes every window with one funnel shift;
//     each window probes a bitmap of the needle codes in shared memory (patterns and, for the '-' strand,
//     reverse(pair(pattern)) -- matched on the forward strand, bigseqkit-lib/locate.go:669-766) without a branch: the
//     16 result bits of a chunk are collected in a mask, and the few set ones (3 % of the windows with 2000 codes) look
//     their 16-bit fingerprint up in a second table; what passes both is parked in a shared-memory queue and confirmed
//     by the whole CTA in an exact table in L2;
//   * header positions and per-tile newline counts are written beside, from which a small second kernel derives the
//     record table (ID slice, sequence length) and turns raw hit positions into (record, 0-based start).
// HBM traffic = N read (+ 4 % look-behind re-read, served by L2) for N algorithmic bytes.
// Outside the grammar (input not FASTA, a header line longer than 960 bytes, more newlines than bases around a lane
// boundary, too many records for the header list) a flag is raised and the caller takes the general path.
#include "kernels.h"
#include "tma.cuh"

namespace bsk {
namespace k {

namespace lt {
constexpr u32 NT = 512;                  // threads per CTA (4 CTAs / SM = 64 warps at 32 registers)
constexpr u32 SPAN = 48;                 // bytes per lane
constexpr u32 NCH = SPAN / 16;           // 16-byte chunks per lane
constexpr u32 REGION = NT * SPAN;        // staged bytes per tile
constexpr u32 LBL = 22;                  // look-behind lanes (all in warp 0)
constexpr u32 LB = LBL * SPAN;           // 1056 look-behind bytes
constexpr u32 T = REGION - LB;           // 23520 owned bytes per tile
constexpr u32 WU = 16;                   // warm-up bytes in front of a lane's span (>= L - 1 symbols unless two newlines fall inside)
constexpr u32 HDR_MAX = 960;             // longest header line accepted (must stay below LB - WU)
constexpr u32 NWARP = NT / 32;
constexpr u32 FBITS = 17;                // first-level bitmap: 2^17 bits = 16 KiB, one probe per window
constexpr u32 FWORDS = (1u << FBITS) / 32;
constexpr u32 PBITS = 12;                // fingerprint table: 2^12 x 16 bits = 8 KiB
constexpr u32 QCAP = 256;                // candidates per tile parked for the confirmation pass
constexpr u32 CTAS = 4;                  // CTAs per SM: one stage buffer each, the other CTAs hide the load
static_assert(T % 16 == 0 && LB % 16 == 0 && SPAN % 16 == 0 && WU % 16 == 0 && HDR_MAX + WU < LB && LBL < 32 && (NCH == 3 || NCH == 6),
              "tile geometry");

// byte classes of the rolling pass
constexpr u8 C_BASE = 8;    // valid base (code in bits 0-1)
constexpr u8 C_RESET = 4;   // ends the run of valid bases (invalid base, header byte)
constexpr u8 C_BREAK = 32;  // header byte: windows never span it
constexpr u8 MARK = 0x01;   // value written over header bytes in the stage buffer

struct Smem {  // dynamic part
  u8 in[REGION + 16];
  u64 full;
  u32 qkey[QCAP], qpos[QCAP], qnl[QCAP];
};
}  // namespace lt

__device__ __forceinline__ u32 lt_nl_flags(u32 w) {
  const u32 x = w ^ 0x0a0a0a0au;
  const u32 y = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
  return ~(y | x) & 0x80808080u;
}

// First level: one bit of a 2^FBITS-bit map per needle code, stored from the top of each word (bit 31 - (index & 31)) so
// that a wrapping left shift by the index puts it into the sign bit of the result.
__device__ __forceinline__ u32 lt_probe(const u32 *filt, u32 h) {
  return __funnelshift_l(0u, filt[h >> (37u - lt::FBITS)], h >> (32u - lt::FBITS));
}
// Second level: open-addressing table of 16-bit fingerprints (0 = empty slot) under a second hash of the code; a window
// that finds its fingerprint is almost surely a needle and goes to the exact table.
__device__ __forceinline__ bool lt_probe2(const unsigned short *fp, u32 h2) {
  u32 slot = h2 >> (32u - lt::PBITS);
  const u32 want = ((h2 >> 5) & 0x7fffu) | 0x8000u;
  for (;;) {
    const u32 e = fp[slot];
    if (e == want) return true;
    if (e == 0u) return false;
    slot = (slot + 1u) & ((1u << lt::PBITS) - 1u);
  }
}

// exact confirmation of a candidate window (code `key`, last base at global byte gend) against the needle table
__device__ __forceinline__ void lt_confirm(const LocateTileArgs &a, u32 key, u32 gend, u32 nl, u32 tile) {
  u32 slot = (key * 0x9E3779B1u) >> a.tshift;
  for (;;) {
    const u32 e = a.table[2 * slot + 1];
    if (e == 0) return;
    if (a.table[2 * slot] == key) {
      for (u32 id = e - 1u; id < a.n_needles && a.nd_code[id] == key; id++) {
        const unsigned long long idx = atomicAdd((unsigned long long *)&a.st->counters[5], 1ull);
        if (idx < a.hit_cap) {
          a.hitA[idx] = ((u64)nl << 32) | (u64)gend;
          a.hitB[idx] = ((u64)tile << 32) | a.nd_ps[id];
        }
      }
      return;
    }
    slot = (slot + 1u) & a.tmask;
  }
}

__global__ void __launch_bounds__(lt::NT, lt::CTAS) k_locate_tile(LocateTileArgs a) {
  using namespace lt;
  BSK_DYN_SMEM(Smem, smp);
  Smem &sm = *smp;
  __align__(16) __shared__ u32 s_filter[FWORDS];
  __align__(16) __shared__ unsigned short s_fp[1u << PBITS];
  __align__(16) __shared__ u8 s_lut[256];
  __shared__ u32 s_wtot[NWARP];
  __shared__ u32 s_decline, s_nl_lb, s_qn;
  const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const u32 n = a.n, n16 = n & ~15u;
  const u32 L = a.L;

  if (tid < 256) s_lut[tid] = a.lut[tid];
  for (u32 i = tid; i < FWORDS; i += NT) s_filter[i] = a.filter[i];
  for (u32 i = tid; i < (1u << PBITS) / 2; i += NT) reinterpret_cast<u32 *>(s_fp)[i] = reinterpret_cast<const u32 *>(a.fptab)[i];
  if (tid == 0) {
    tma::mbar_init(&sm.full, 1);
    tma::fence_barrier_init();
  }
  __syncthreads();

  // region of tile t = global bytes [t*T - LB, t*T + T); the bulk copy brings the whole 16-byte chunks inside the file
  auto bulk_range = [&](u32 tile, u32 &g0, u32 &g1) {
    const u32 t0 = tile * T;
    g0 = t0 >= LB ? t0 - LB : 0u;
    g1 = t0 + T < n16 ? t0 + T : n16;
    return g1 > g0;
  };
  auto issue = [&](u32 tile) {
    u32 g0, g1;
    if (bulk_range(tile, g0, g1)) {
      tma::mbar_expect_tx(&sm.full, g1 - g0);
      tma::bulk_load(&sm.in[g0 + LB - tile * T], a.in + g0, g1 - g0, &sm.full);
    }
  };
  if (tid == 0 && blockIdx.x < a.n_tiles) issue(blockIdx.x);

  const u32 kmul = a.kmul, kmul2 = a.kmul2;
  u32 it = 0;
  for (u32 tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, it++) {
    const u32 parity = it & 1u;
    const u32 t0 = tile * T;
    u8 *d = sm.in;                                           // region byte i == global byte t0 - LB + i
    const u32 lim = (n - t0 < T ? n - t0 : T) + LB;          // valid bytes of the region (look-behind included)
    {
      u32 g0, g1;
      if (bulk_range(tile, g0, g1)) tma::mbar_wait(&sm.full, parity);
    }
    // bytes the bulk copy did not bring: in front of the file (tile 0), the ragged tail, '\n' padding
    if (t0 < LB || t0 + T > n16) {
      for (u32 i = tid; i < REGION; i += NT) {
        const bool before = t0 + i < LB;
        const u32 g = t0 + i - LB;
        if (before || g >= n16) d[i] = (!before && g < n) ? a.in[g] : (u8)'\n';
      }
      __syncthreads();
    }
    if (tid == 0) { s_decline = 0; s_qn = 0; }

    // ---- newline masks of this lane's span (bit j of m[k]: byte 32 k + j is '\n'), CTA-wide prefix
    const u32 span0 = tid * SPAN;
    u32 m[3];
    {
      u32 m16[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (u32 j = 0; j < NCH; j++) {
        const uint4 v = *reinterpret_cast<const uint4 *>(d + span0 + j * 16u);
        u32 lo = __dp4a(lt_nl_flags(v.x), 0x08040201u, 0u);
        lo = __dp4a(lt_nl_flags(v.y), 0x80402010u, lo);
        u32 hi = __dp4a(lt_nl_flags(v.z), 0x08040201u, 0u);
        hi = __dp4a(lt_nl_flags(v.w), 0x80402010u, hi);
        m16[j] = (lo >> 7) | (hi << 1);
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
    const u32 cnt = (u32)(__popc(m[0]) + __popc(m[1]) + __popc(m[2]));
    u32 inc = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const u32 y = __shfl_up_sync(0xffffffffu, inc, off);
      if ((int)lane >= off) inc += y;
    }
    if (lane == 31) s_wtot[warp] = inc;
    if (tid == LBL) s_nl_lb = inc - cnt;  // newlines in the look-behind = exclusive prefix of lane LBL (warp 0)
    __syncthreads();
    u32 nl_before = inc - cnt;  // newlines of the region in front of this lane's span
    u32 nl_total = 0;
#pragma unroll
    for (u32 w = 0; w < NWARP; w++) {
      const u32 x = s_wtot[w];
      if (w < warp) nl_before += x;
      nl_total += x;
    }
    const u32 nl_lookbehind = s_nl_lb;
    if (tid == 0) a.tile_nl[tile] = nl_total - nl_lookbehind;

    // ---- header lines: '>' right after a newline.  The lane that owns the newline blanks the header's bytes.
    {
#pragma unroll
      for (u32 k2 = 0; k2 < 3; k2++) {
        u32 mm = m[k2];
        while (mm) {
          const u32 b = (u32)__ffs((int)mm) - 1u;
          mm &= mm - 1u;
          const u32 x = span0 + k2 * 32u + b;  // region position of the newline
          const u32 h = x + 1u;
          if (h < lim && d[h] == '>') {
            u32 j = h;
            while (j < lim && d[j] != '\n') {
              d[j] = MARK;
              j++;
              if (j - h > HDR_MAX) { s_decline = 1; break; }
            }
            if (h >= LB) {  // starts inside the owned range: one entry of the record list
              // newlines in [t0, h): everything up to and including the newline at x, minus the look-behind's
              u32 upto = nl_before;
              for (u32 q = 0; q < k2; q++) upto += (u32)__popc(m[q]);
              upto += (u32)__popc(m[k2] & ((2u << b) - 1u));
              const unsigned long long idx = atomicAdd((unsigned long long *)&a.st->counters[6], 1ull);
              if (idx < a.hdr_cap) {
                a.hdr_off[idx] = (u64)t0 + (h - LB);
                a.hdr_nl[idx] = (u64)(upto - nl_lookbehind);
              }
            }
          }
        }
      }
      if (tile == 0 && tid == 0 && n > 0 && a.in[0] != '>') s_decline = 1;  // the input must open with a record
    }
    __syncthreads();

    // ---- 2-bit codes of the lane's span, one 16-byte chunk at a time (owned lanes only).  A chunk of 16 plain bases,
    // or of 15 and one newline, is packed with SWAR arithmetic (first base in the highest bits) and appended to the
    // codes of the 16 bases before it; every window is then one funnel shift away.  Each window probes the
    // bitmap; the few that also find their fingerprint are parked in the queue and confirmed after the loop.  Any other chunk
    // (invalid bases, header bytes, several newlines) goes byte by byte through the class table.
    // newline flags of the 16 bytes in front of the span = of the previous lane's last chunk
    const u32 prev_m2 = __shfl_up_sync(0xffffffffu, NCH == 3 ? (m[1] << 16) : m[2], 1);
    if (tid >= LBL && span0 < lim) {
      u32 code = 0, run = 0;  // codes of the last 16 bases, length of the run of valid bases behind them (saturating)
      const u32 vmask = a.vmask, vbase = a.vbase;
      // hit: window whose last base is byte `rel` of the span
      auto park = [&](u32 c, u32 rel) {
        const u32 p = span0 + rel;  // region position of the window's last base
        if (p >= lim) return;
        u32 nls = 0;  // newlines of the span in front of p
        if (rel >= 32u) nls += (u32)__popc(m[0]);
        if (rel >= 64u) nls += (u32)__popc(m[1]);
        const u32 mk = rel < 32u ? m[0] : (rel < 64u ? m[1] : m[2]);
        nls += (u32)__popc(mk & ((1u << (rel & 31u)) - 1u));
        const u32 qi = atomicAdd(&s_qn, 1u);
        const u32 key = c & a.cmask, gend = t0 + (p - LB), nl = nl_before - nl_lookbehind + nls;
        if (qi < QCAP) { sm.qkey[qi] = key; sm.qpos[qi] = gend; sm.qnl[qi] = nl; }
        else lt_confirm(a, key, gend, nl, tile);  // queue full (a panel that matches everywhere): confirm in place
      };
      // byte-wise path over one chunk at offset rel0 of the span (the warm-up does not probe)
      auto bytewise = [&](const uint4 &q, u32 rel0, bool probe, u32 &nb, u32 &brk) {
#pragma unroll 1
        for (u32 i = 0; i < 4; i++) {
          const u32 w = i == 0 ? q.x : (i == 1 ? q.y : (i == 2 ? q.z : q.w));
#pragma unroll 1
          for (u32 j = 0; j < 4; j++) {
            const u32 c = s_lut[(w >> (8u * j)) & 0xffu];
            if (c & C_BASE) {
              code = code * 4u + (c & 3u);
              run++;
              if (probe) {
                const u32 h1 = code * kmul;  // depends on the last L bases only (kmul = odd << (32 - 2L))
                if ((int)lt_probe(s_filter, h1) < 0 && run >= L && lt_probe2(s_fp, code * kmul2)) park(code, rel0 + 4u * i + j);
              }
            }
            if (c & C_RESET) run = 0;
            nb += (c >> 4) & 1u;
            brk |= c;
          }
        }
      };
      // packs the chunk; true when it is 16 bases (cnt = 16) or 15 bases and the newline at byte k (cnt = 15)
      auto pack = [&](const uint4 &q, u32 nl16, u32 &lo, u32 &hi, u32 &cnt, u32 &k) {
        const u32 one = (nl16 != 0u && (nl16 & (nl16 - 1u)) == 0u) ? 1u : 0u;
        k = nl16 ? (u32)__ffs((int)nl16) - 1u : 16u;
        // a single newline is turned into the first base letter before the test and its (zero) code removed afterwards
        const u32 fix = one ? ((0x0au ^ (vbase & 0xffu)) << (8u * (k & 3u))) : 0u;
        const u32 kw = one ? (k >> 2) : 4u;
        u32 w4[4] = {q.x, q.y, q.z, q.w};
        u32 bad = 0, p = 0;
#pragma unroll
        for (u32 i = 0; i < 4; i++) {
          const u32 wm = (w4[i] ^ (kw == i ? fix : 0u)) & vmask;
          const u32 x = (wm >> 1) & 0x03030303u;                 // A 0, C 1, T 2, G 3 (either case)
          const u32 t = (x >> 1) & ~x & 0x01010101u;             // the T bytes
          const u32 expect = x * 2u + vbase + t * 0x0fu;         // the letter each code stands for: 41 43 47, 54 = 45 + 0f
          bad |= wm ^ expect;
          p = (p << 8) | ((x * 0x40100401u) >> 24);              // four codes -> one byte, first base highest
        }
        if (bad != 0u || (nl16 != 0u && !one)) return false;
        if (one) {
          const u32 below = 0x3fffffffu >> (2u * k);             // the codes behind the newline
          const u32 p30 = ((p & ~(0xffffffffu >> (2u * k))) >> 2) | (p & below);
          lo = (code << 30) | p30;
          hi = code >> 2;
          cnt = 15u;
        } else {
          lo = p;
          hi = code;
          cnt = 16u;
        }
        return true;
      };
      {  // warm-up: the WU bytes in front of the span (no probes)
        const uint4 q = *reinterpret_cast<const uint4 *>(d + span0 - WU);
        u32 nlw;
        if (lane != 0u) nlw = prev_m2 >> 16;
        else {
          u32 lo2 = __dp4a(lt_nl_flags(q.x), 0x08040201u, 0u);
          lo2 = __dp4a(lt_nl_flags(q.y), 0x80402010u, lo2);
          u32 hi2 = __dp4a(lt_nl_flags(q.z), 0x08040201u, 0u);
          hi2 = __dp4a(lt_nl_flags(q.w), 0x80402010u, hi2);
          nlw = (lo2 >> 7) | (hi2 << 1);
        }
        u32 lo, hi, cnt, k;
        if (pack(q, nlw, lo, hi, cnt, k)) {
          code = lo;
          run = cnt;
        } else {
          u32 nb = 0, brk = 0;
          bytewise(q, 0u, false, nb, brk);
          // too few symbols in the warm-up to judge the first windows of the span (a run of > 16 newlines)
          if (nb + 1u < L && !(brk & C_BREAK) && t0 + span0 >= LB + WU) s_decline = 1;
        }
      }
      const uint4 *vp = reinterpret_cast<const uint4 *>(d + span0);
#pragma unroll 1
      for (u32 v = 0; v < NCH; v++) {
        const uint4 q = vp[v];
        const u32 mw = (v >> 1) == 0u ? m[0] : ((v >> 1) == 1u ? m[1] : m[2]);
        const u32 nl16 = (v & 1u) ? (mw >> 16) : (mw & 0xffffu);
        u32 lo, hi, cnt, k;
        if (pack(q, nl16, lo, hi, cnt, k)) {
          const u32 need = L > run ? L - run : 0u;  // windows whose last base is base >= need - 1 of the chunk are whole
          u32 hm = 0;                               // windows that pass the first probe (bit 15 - i)
#pragma unroll
          for (u32 i = 0; i < 16; i++) {            // i = bases of the chunk behind the window's last one
            const u32 c = __funnelshift_r(lo, hi, 2u * i);
            hm = __funnelshift_l(lt_probe(s_filter, c * kmul), hm, 1u);  // the sign bit enters from below
          }
          if (cnt != 16u) hm &= ~1u;                // 15 bases: the window 15 bases back ended in the previous chunk
          while (hm) {                              // a few per warp and chunk
            const u32 i = 16u - (u32)__ffs((int)hm);
            hm &= hm - 1u;
            const u32 c = __funnelshift_r(lo, hi, 2u * i);
            const u32 jb = cnt - 1u - i;            // index of the window's last base among the chunk's bases
            if (jb + 1u >= need && lt_probe2(s_fp, c * kmul2)) park(c, v * 16u + jb + (jb >= k ? 1u : 0u));
          }
          code = lo;
          run = run + cnt;
          if (run > 64u) run = 64u;
        } else {
          u32 nb = 0, brk = 0;
          bytewise(q, v * 16u, true, nb, brk);
        }
      }
    }
    __syncthreads();  // every thread is done with the stage; the queue is complete
    if (tid == 0) {
      if (s_decline) atomicAdd((unsigned long long *)&a.st->counters[0], 1ull);
      const u32 tn = tile + gridDim.x;
      if (tn < a.n_tiles) issue(tn);  // refill the stage while the candidates are confirmed
    }
    {
      const u32 qn = s_qn < QCAP ? s_qn : QCAP;
      for (u32 i = tid; i < qn; i += NT) lt_confirm(a, sm.qkey[i], sm.qpos[i], sm.qnl[i], tile);
    }
    __syncthreads();  // s_qn / the queue are read by everybody before thread 0 clears them for the next tile
  }
}

// record table from the sorted header list: one thread per record
__global__ void k_locate_records(const u8 *__restrict__ in, u32 n, const u64 *__restrict__ hdr_off, const u64 *__restrict__ hdr_nl,
                                 const u32 *__restrict__ tile_nl_base, u32 n_tiles, u32 n_rec, u32 *name_off, u32 *name_len,
                                 u32 *seq_start, u32 *seq_nl, u32 *seq_len) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rec) return;
  const u32 h = (u32)hdr_off[r];
  const u32 nlg_h = tile_nl_base[h / lt::T] + (u32)hdr_nl[r];  // newlines in [0, h)
  u32 e = h + 1;
  while (e < n && in[e] != '\n') e++;  // header line (the bytes after '>')
  name_off[r] = h + 1;
  name_len[r] = e - (h + 1);
  const u32 s0 = e < n ? e + 1 : n;     // first sequence byte
  const u32 nl0 = nlg_h + (e < n ? 1u : 0u);  // newlines in [0, s0)
  u32 next = n, nl_next = tile_nl_base[n_tiles];
  if (r + 1 < n_rec) {
    next = (u32)hdr_off[r + 1];
    nl_next = tile_nl_base[next / lt::T] + (u32)hdr_nl[r + 1];
  }
  seq_start[r] = s0;
  seq_nl[r] = nl0;
  seq_len[r] = next > s0 ? (next - s0) - (nl_next - nl0) : 0u;
}

// raw hit (position of the window's last base, newlines of its tile in front of it) -> the keys the row sorter wants:
// A = record << 32 | pattern << 1 | strand, B = coordinate on the strand << 32 | 0-based start on the '+' strand
__global__ void k_locate_resolve(u64 *hitA, u64 *hitB, u64 n_hits, const u64 *__restrict__ hdr_off, u32 n_rec,
                                 const u32 *__restrict__ tile_nl_base, const u32 *__restrict__ seq_start,
                                 const u32 *__restrict__ seq_nl, const u32 *__restrict__ seq_len, u32 L) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_hits) return;
  const u32 gend = (u32)hitA[i], nl_local = (u32)(hitA[i] >> 32);
  const u32 tile = (u32)(hitB[i] >> 32), ps = (u32)hitB[i];
  u32 lo = 0, hi = n_rec;  // last record whose header starts at or before gend
  while (hi - lo > 1) {
    const u32 mid = lo + ((hi - lo) >> 1);
    if ((u32)hdr_off[mid] <= gend) lo = mid;
    else hi = mid;
  }
  const u32 r = lo;
  const u32 nlg = tile_nl_base[tile] + nl_local;
  const u32 e = (gend - seq_start[r]) - (nlg - seq_nl[r]);  // 0-based index of the window's last base in the record
  const u32 q = e + 1u - L;
  const u32 l = seq_len[r];
  const u32 coord = (ps & 1u) ? l - q - L : q;
  hitA[i] = ((u64)r << 32) | ps;
  hitB[i] = ((u64)coord << 32) | q;
}

u32 locate_tile_tiles(u32 n) { return (n + lt::T - 1) / lt::T; }
u32 locate_tile_bytes() { return lt::T; }
u32 locate_tile_filter_bits() { return lt::FBITS; }
u32 locate_tile_fp_bits() { return lt::PBITS; }

void locate_tile(LocateTileArgs a, int n_sm, cudaStream_t s) {
  a.n_tiles = locate_tile_tiles(a.n);
  const size_t smem = sizeof(lt::Smem) + 16;
#ifndef BSK_EMU
  static size_t attr_set[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && attr_set[dev] < smem) {
    cudaFuncSetAttribute(k_locate_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set[dev] = smem;
  }
#endif
  u32 grid = (u32)n_sm * lt::CTAS;
  if (grid > a.n_tiles) grid = a.n_tiles;
  if (grid == 0) return;
  BSK_LAUNCH(k_locate_tile, grid, lt::NT, smem, s, a);
}

void locate_records(const u8 *in, u32 n, const u64 *hdr_off, const u64 *hdr_nl, const u32 *tile_nl_base, u32 n_tiles, u32 n_rec,
                    u32 *name_off, u32 *name_len, u32 *seq_start, u32 *seq_nl, u32 *seq_len, cudaStream_t s) {
  if (!n_rec) return;
  BSK_LAUNCH_FLAT(k_locate_records, (n_rec + 127) / 128, 128, 0, s, in, n, hdr_off, hdr_nl, tile_nl_base, n_tiles, n_rec, name_off,
                  name_len, seq_start, seq_nl, seq_len);
}

void locate_resolve(u64 *hitA, u64 *hitB, u64 n_hits, const u64 *hdr_off, u32 n_rec, const u32 *tile_nl_base, const u32 *seq_start,
                    const u32 *seq_nl, const u32 *seq_len, u32 L, cudaStream_t s) {
  if (!n_hits) return;
  BSK_LAUNCH_FLAT(k_locate_resolve, (u32)((n_hits + 255) / 256), 256, 0, s, hitA, hitB, n_hits, hdr_off, n_rec, tile_nl_base,
                  seq_start, seq_nl, seq_len, L);
}

}  // namespace k
}  // namespace bsk
