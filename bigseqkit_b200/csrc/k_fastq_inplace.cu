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
#include "kernels.h"
#include "tma.cuh"

namespace bsk {
namespace k {

namespace fq {
constexpr u32 T = 20480;       // tile bytes
constexpr u32 H = 4096;        // halo bytes
constexpr u32 PRE = 16;        // look-behind bytes in front of the tile
constexpr u32 NT = 512;        // threads per CTA
constexpr u32 NWARP = NT / 32;
constexpr u32 CPL = 3;         // 16-byte chunks per lane in the newline scan
constexpr u32 STAGE = PRE + T + H + 16;
constexpr u32 NSTAGE = 3;
constexpr u32 LCAP = 3072;     // line starts per region
constexpr u32 RCAP = 512;      // owned records per tile (slot stride)
static_assert((T + H) / 16 == NWARP * 32 * CPL, "scan partition must cover the region exactly");
static_assert(STAGE % 16 == 0, "stage size");

struct Smem {
  u8 in[NSTAGE][STAGE];
  u64 full[NSTAGE];
  u8 lut[256];
  u16 ls[LCAP + 8];     // line starts, ls[0] = 0
  u8 isrs[LCAP + 8];    // line k opens a record
  u16 r_line[RCAP];     // first line of every owned record, in input order
  u32 wtot[NWARP];
  u32 wtot2[NWARP];
  u32 bad;
};
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

// In-place rewrite of the byte range [a, a + L) of the region (d = region byte 0, 4-byte aligned) by a group
// of G lanes: optional reversal, optional byte map.  Whole 32-bit words inside the range are produced with one
// PRMT from two source words (registers hold the segment until every lane has read); the <= 3 ragged bytes at
// each end go through a byte path on lanes 0-5 of the group.  Every lane of the warp must call it.
template <bool REV, bool LUT, u32 G, u32 WPL>
__device__ __forceinline__ void seg_inplace(u8 *d, u32 a, u32 L, const u8 *lut, u32 gl) {
  u32 *w32 = reinterpret_cast<u32 *>(d);
  const u32 e = a + L;
  const u32 ai = (a + 3u) & ~3u, ae = e & ~3u;          // word-aligned interior [ai, ae)
  const u32 nwf = ae > ai ? (ae - ai) >> 2 : 0u;
  const bool slow = __any_sync(0xffffffffu, nwf > G * WPL);
  if (!slow) {
    // ragged bytes: lanes 0-2 the head [a, min(ai, e)), lanes 3-5 the tail [max(ae, ai), e)
    const u32 head_end = ai < e ? ai : e;
    const u32 tail_beg = ae > ai ? ae : (ai < e ? ai : e);
    u32 x = gl < 3u ? a + gl : tail_beg + (gl - 3u);
    const bool eok = gl < 3u ? x < head_end : (gl < 6u && x < e);
    u8 ev = 0;
    if (eok) {
      ev = d[REV ? (a + e - 1u - x) : x];
      if (LUT) ev = lut[ev];
    }
    u32 vals[WPL];
    const u32 U0 = a + e - 4u - ai;  // lowest source byte of interior word 0 (dest byte ai+b <- source U0+3-b)
    const u32 sh = U0 & 3u;
    const u32 sel = (sh + 3u) | ((sh + 2u) << 4) | ((sh + 1u) << 8) | (sh << 12);
    const u32 q0 = U0 >> 2, wi = ai >> 2;
#pragma unroll
    for (u32 j = 0; j < WPL; j++) {
      const u32 idx = gl + j * G;
      u32 v = 0;
      if (idx < nwf) {
        if (REV) {
          const u32 lo = w32[q0 - idx], hi = w32[q0 - idx + 1u];
          v = __byte_perm(lo, hi, sel);
        } else {
          v = w32[wi + idx];
        }
        if (LUT) v = lut4(lut, v);
      }
      vals[j] = v;
    }
    __syncwarp();
#pragma unroll
    for (u32 j = 0; j < WPL; j++) {
      const u32 idx = gl + j * G;
      if (idx < nwf) w32[wi + idx] = vals[j];
    }
    if (eok) d[x] = ev;
    __syncwarp();
  } else {
    // long segment: independent byte pairs (i, L-1-i), no hazards
    if (REV) {
      const u32 half = L >> 1;
      for (u32 i = gl; i < half; i += G) {
        u8 x = d[a + i], y = d[a + L - 1 - i];
        if (LUT) { x = lut[x]; y = lut[y]; }
        d[a + i] = y;
        d[a + L - 1 - i] = x;
      }
      if (LUT && (L & 1u) && gl == 0) d[a + half] = lut[d[a + half]];
    } else if (LUT) {
      for (u32 i = gl; i < L; i += G) d[a + i] = lut[d[a + i]];
    }
    __syncwarp();
  }
}

// transform of all owned records of a tile, G lanes per record
template <u32 G, u32 WPL>
__device__ __forceinline__ void transform_tile(fq::Smem &sm, u8 *d, u32 n_own, int reverse, int use_lut) {
  const u32 g = threadIdx.x / G, gl = threadIdx.x % G;
  for (u32 rb = 0; rb < n_own; rb += fq::NT / G) {  // uniform trip count per CTA
    const u32 r = rb + g;
    u32 so = 0, sl = 0, qo = 0;
    if (r < n_own) {
      const u32 k = sm.r_line[r];
      so = sm.ls[k + 1];
      sl = sm.ls[k + 2] - 1u - so;
      qo = sm.ls[k + 3];
    }
    if (reverse) {
      if (use_lut) seg_inplace<true, true, G, WPL>(d, so, sl, sm.lut, gl);
      else seg_inplace<true, false, G, WPL>(d, so, sl, sm.lut, gl);
      seg_inplace<true, false, G, WPL>(d, qo, sl, sm.lut, gl);
    } else {
      seg_inplace<false, true, G, WPL>(d, so, sl, sm.lut, gl);
    }
  }
}

__global__ void __launch_bounds__(fq::NT, 2) k_fastq_inplace(FqInplaceArgs a) {
  using namespace fq;
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

  // bulk load of the region of tile `tile` into stage s; returns nothing, bytes may be 0 (then no barrier phase)
  auto issue = [&](u32 tile, u32 s) {
    const long long r0 = (long long)tile * T - PRE;  // global position of stage byte 0
    long long g0 = r0 < 0 ? 0 : r0;
    long long g1 = (long long)tile * T + T + H;
    if (g1 > (long long)n16) g1 = n16;
    if (g1 > g0) {
      const u32 bytes = (u32)(g1 - g0);
      tma::mbar_expect_tx(&sm.full[s], bytes);
      tma::bulk_load(&sm.in[s][(u32)(g0 - r0)], a.in + g0, bytes, &sm.full[s]);
    }
  };
  auto has_bulk = [&](u32 tile) {
    const long long r0 = (long long)tile * T - PRE;
    long long g0 = r0 < 0 ? 0 : r0;
    long long g1 = (long long)tile * T + T + H;
    if (g1 > (long long)n16) g1 = n16;
    return g1 > g0;
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
    if (tid == 0) {
      const u32 tn = tile + (NSTAGE - 1) * gridDim.x;
      if (it > 0) tma::bulk_wait_read();  // the stage being refilled was the source of the previous tile's store
      if (tn < a.n_tiles) issue(tn, (it + NSTAGE - 1) % NSTAGE);
    }
    const u32 t0 = tile * T;
    u8 *stage = sm.in[s];
    u8 *d = stage + PRE;  // region byte 0 == global byte t0
    const u32 lim = (n - t0 < T + H) ? n - t0 : T + H;  // valid bytes of the region
    const bool eof = (n - t0) <= T + H;
    if (has_bulk(tile)) tma::mbar_wait(&sm.full[s], parity);
    // bytes the bulk copy did not bring: the look-behind of tile 0, the ragged tail of the file, padding
    if (tile == 0 || (unsigned long long)t0 + T + H > n16) {
      const long long r0 = (long long)t0 - PRE;
      for (u32 i = tid; i < PRE + T + H + 16; i += NT) {
        const long long g = r0 + i;
        if (g < 0 || g >= (long long)n16) stage[i] = (g >= 0 && g < (long long)n) ? a.in[g] : (u8)'\n';
      }
      __syncthreads();
    }

    // ---- newline scan: lane owns CPL consecutive 16-byte chunks; the 0x80 flag bytes of a chunk are packed into a
    // position-ordered 16-bit mask with four IDP.4A (flag byte * {1,2,4,8} summed = nibble << 7)
    const u32 span = (warp * 32u + lane) * (CPL * 16u);
    u32 mlo, mhi;
    {
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
    }
    if (span + CPL * 16u > lim) {  // bytes past the end of the file do not count
      const u32 valid = lim > span ? lim - span : 0u;
      if (valid < 32u) { mlo &= (1u << valid) - 1u; mhi = 0; }
      else mhi &= (1u << (valid - 32u)) - 1u;
    }
    const u32 cnt = __popc(mlo) + __popc(mhi);
    u32 inc = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const u32 y = __shfl_up_sync(0xffffffffu, inc, off);
      if ((int)lane >= off) inc += y;
    }
    if (lane == 31) sm.wtot[warp] = inc;
    if (tid == 0) sm.bad = 0;
    __syncthreads();
    u32 base, n_nl;
    {
      u32 x = sm.wtot[lane & (NWARP - 1u)];
#pragma unroll
      for (int off = 1; off < (int)NWARP; off <<= 1) {
        const u32 y = __shfl_up_sync(0xffffffffu, x, off, NWARP);
        if ((int)(lane & (NWARP - 1u)) >= off) x += y;
      }
      n_nl = __shfl_sync(0xffffffffu, x, NWARP - 1u);
      base = __shfl_sync(0xffffffffu, x, (warp + NWARP - 1u) & (NWARP - 1u));
      if (warp == 0) base = 0;
    }
    const bool virt = eof && lim > 0 && d[lim - 1] != '\n';  // unterminated last line
    const bool overflow = n_nl + 2 > LCAP;
    if (!overflow) {
      u32 k = base + inc - cnt + 1;  // ls[k] = start of the line after the (k-1)-th newline
      while (mlo) {
        const u32 t = (u32)__ffs((int)mlo) - 1u;
        mlo &= mlo - 1u;
        sm.ls[k++] = (u16)(span + t + 1u);
      }
      while (mhi) {
        const u32 t = (u32)__ffs((int)mhi) - 1u;
        mhi &= mhi - 1u;
        sm.ls[k++] = (u16)(span + 32u + t + 1u);
      }
      if (tid == 0) {
        sm.ls[0] = 0;
        if (virt) sm.ls[n_nl + 1] = (u16)(lim + 1);
      }
    }
    if (virt) n_nl++;
    __syncthreads();
    if (overflow) {
      if (tid == 0) atomicAdd((unsigned long long *)&a.st->counters[0], 1ull);
      continue;  // uniform
    }

    // ---- record starts: line k (k <= n_nl) opens a record iff it starts with '@' and the line before is not a bare "+"
    // whose own predecessor ended ... (rule pinned in SURVEY C.1: '\n@' unless preceded by "\n+")
    for (u32 k = tid; k <= n_nl; k += NT) {
      const u32 p = sm.ls[k];
      bool rs = p < lim && d[p] == '@';
      if (rs && k == 0) rs = (tile == 0) || d[-1] == '\n';
      if (rs && (unsigned long long)t0 + p >= 3ull && d[(int)p - 3] == '\n' && d[(int)p - 2] == '+' && !(k == 0 && tile == 0))
        rs = false;
      sm.isrs[k] = rs ? 1 : 0;
    }
    __syncthreads();

    // ---- owned records (start inside the tile), grammar check, ordered compaction
    u32 run = 0;
    for (u32 kb = 0; kb <= n_nl; kb += NT) {  // uniform trip count
      const u32 k = kb + tid;
      bool own = false;
      if (k <= n_nl && sm.isrs[k] && sm.ls[k] < T) {
        own = true;
        bool ok = k + 4 <= n_nl;  // the four newlines of the record are inside the region
        if (ok) {
          const u32 l1 = sm.ls[k + 1], l2 = sm.ls[k + 2], l3 = sm.ls[k + 3], l4 = sm.ls[k + 4];
          const u32 sl = l2 - 1 - l1, ql = l4 - 1 - l3;
          ok = !sm.isrs[k + 1] && !sm.isrs[k + 2] && !sm.isrs[k + 3];
          ok = ok && (l3 - l2 == 2) && d[l2] == '+';         // bare "+" line
          ok = ok && sl == ql && !(sl > 0 && d[l1] == '+');  // a sequence line starting with '+' flips the parser
          ok = ok && (sm.isrs[k + 4] || (eof && l4 >= lim)); // next line opens a record, or the file ends here
        }
        if (!ok) sm.bad = 1;
      }
      const u32 bal = __ballot_sync(0xffffffffu, own);
      if (lane == 0) sm.wtot2[warp] = __popc(bal);
      __syncthreads();
      u32 wb, tot;
      {
        u32 x = sm.wtot2[lane & (NWARP - 1u)];
#pragma unroll
        for (int off = 1; off < (int)NWARP; off <<= 1) {
          const u32 y = __shfl_up_sync(0xffffffffu, x, off, NWARP);
          if ((int)(lane & (NWARP - 1u)) >= off) x += y;
        }
        tot = __shfl_sync(0xffffffffu, x, NWARP - 1u);
        wb = __shfl_sync(0xffffffffu, x, (warp + NWARP - 1u) & (NWARP - 1u));
        if (warp == 0) wb = 0;
      }
      if (own) {
        const u32 r = run + wb + __popc(bal & ((1u << lane) - 1u));
        if (r < RCAP) sm.r_line[r] = (u16)k;
      }
      run += tot;
      __syncthreads();
    }
    const u32 n_own = run;
    bool bad = sm.bad != 0 || n_own > RCAP;
    if (tile == 0 && !sm.isrs[0]) bad = true;  // the file must open with a marked record
    if (bad) {
      if (tid == 0) atomicAdd((unsigned long long *)&a.st->counters[0], 1ull);
      __syncthreads();
      continue;
    }

    // ---- element slots + in-place transform
    for (u32 r = tid; r < n_own; r += NT) a.slots[(size_t)tile * RCAP + r] = sm.ls[sm.r_line[r]];
    if (tid == 0) a.tile_cnt[tile] = n_own;
    if (a.reverse || a.use_lut) {
      if (a.group == 8) transform_tile<8, 8>(sm, d, n_own, a.reverse, a.use_lut);
      else if (a.group == 16) transform_tile<16, 4>(sm, d, n_own, a.reverse, a.use_lut);
      else transform_tile<32, 4>(sm, d, n_own, a.reverse, a.use_lut);
    }
    tma::fence_proxy_async();
    __syncthreads();

    // ---- owned byte range [lo, hi) -> out, same offsets: bulk store of the aligned body, ragged ends by warp 0
    if (warp == 0 && n_own > 0) {
      const u32 lo = sm.ls[sm.r_line[0]];
      const u32 kl = sm.r_line[n_own - 1];
      const u32 hi = sm.ls[kl + 4];  // start of the next record == one past the '\n' that ends the last owned one
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
    // no barrier here: warp 0 reaches the next tile's barriers only after it has read what it needs
  }
  if (tid == 0) tma::bulk_wait_all();
}

// element offsets from the per-tile slot lists: one warp per tile
__global__ void k_fastq_elem_expand(const u32 *__restrict__ tile_cnt, const u64 *__restrict__ tile_base,
                                    const u16 *__restrict__ slots, u64 *__restrict__ elem_off, u32 n_tiles) {
  const u32 tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (tile >= n_tiles) return;
  const u32 c = tile_cnt[tile];
  const u64 b = tile_base[tile];
  for (u32 r = lane; r < c; r += 32) elem_off[b + r] = (u64)tile * fq::T + slots[(size_t)tile * fq::RCAP + r];
}

u32 fastq_inplace_tiles(u32 n) { return (n + fq::T - 1) / fq::T; }
u32 fastq_inplace_slot_stride() { return fq::RCAP; }

void fastq_inplace(const u8 *in, u32 n, u8 *out, const u8 *lut, u32 *tile_cnt, u16 *slots, DevStatus *st, int reverse,
                   int use_lut, int group, int n_sm, cudaStream_t s) {
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
  a.group = group;
  const size_t smem = sizeof(fq::Smem) + 16;
#ifndef BSK_EMU
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_fastq_inplace, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
#endif
  u32 grid = (u32)n_sm * 2u;
  if (grid > a.n_tiles) grid = a.n_tiles;
  if (grid == 0) return;
  BSK_LAUNCH(k_fastq_inplace, grid, fq::NT, smem, s, a);
}

void fastq_elem_expand(const u32 *tile_cnt, const u64 *tile_base, const u16 *slots, u64 *elem_off, u32 n_tiles,
                       cudaStream_t s) {
  if (!n_tiles) return;
  const u64 threads = (u64)n_tiles * 32;
  BSK_LAUNCH_FLAT(k_fastq_elem_expand, (u32)((threads + 255) / 256), 256, 0, s, tile_cnt, tile_base, slots, elem_off, n_tiles);
}

}  // namespace k
}  // namespace bsk
