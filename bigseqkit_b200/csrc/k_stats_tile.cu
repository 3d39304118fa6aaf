// k_stats_tile.cu -- `stats` on short records (reads, single- or multi-line FASTA) in ONE streaming pass.
//
//   Stats.Call        bigseqkit-lib/stats.go:48-117   hist[len(seq)]++ ; with -a: Q20/Q30 per quality byte
//                                                     (:90-100), gap letters per sequence byte (:102)
//   SeqParser.Read    bigseqkit-lib/helper.go:219-325  (which bytes are sequence / quality)
//
// Same skeleton as k_fastq_inplace.cu (tile_common.cuh): persistent CTAs, TMA bulk loads into a 2-stage ring,
// newline scan -> line starts with record flags in shared memory.  Every owned record (it STARTS inside the tile)
// adds its sequence length to a per-CTA shared-memory histogram; with -a the sequence and quality lines become
// work items whose bytes are counted 4 at a time (SWAR) by groups of 8 lanes.  Nothing is written but a 32 KiB
// histogram and three counters, so the kernel moves N bytes for N algorithmic bytes.
// Anything outside the short-record grammar (FASTQ that is not 4-line, records longer than the halo, a file that does
// not open with a marker) raises a flag and the caller takes the general path.
#include "kernels.h"
#include "tile_common.cuh"

namespace bsk {
namespace k {

namespace st {
constexpr u32 HB = 4224;  // histogram bins: a record (and so its sequence) is shorter than the halo + slack
typedef tile::Geo<512, 3, 3, 2, 2048> G;
constexpr u32 ICAP = G::LCAP;  // work items (sequence / quality lines) per tile

// The shared-memory histogram holds the lengths below HBS (reads); longer ones go straight to the global histogram.
// Without -a 4 CTAs / SM fit (64 warps at 32 registers); -a needs the work-item list (8 KiB): 3 CTAs / SM.
constexpr u32 HBS = 960;
template <bool ALL>
struct Smem {
  u8 in[G::NSTAGE][G::STAGE];
  u64 full[G::NSTAGE];
  u16 ls[G::LCAP + 8];
  u32 item[ALL ? ICAP : 1];  // a (15 bits) | len (15 bits) << 15 | kind << 30   (kind 1 = quality line)
  u32 hist[HBS];
  u32 wtot[G::NWARP];
  u32 bad, rescan, n_item, n_rec;
  unsigned long long acc[3];  // q20, q30, gaps of this CTA
};
}  // namespace st

struct StatsTileArgs {
  const u8 *in;
  u32 n;
  u64 *hist;        // [HB] dense length histogram (global)
  DevStatus *st;    // counters: [0] declined tiles, [1] q20, [2] q30, [3] gap letters, [5] records
  u32 n_tiles;
  int fastq, all, fq_offset;
  u32 gap[4];       // up to four gap letters, each replicated into 4 bytes; n_gap of them are valid
  int n_gap;
  int gap_below_40; // every gap letter is below '@' (0x40) and '@' itself is not one: whole words can be skipped
  u32 scan_halo;
};

// flags (0x80 per byte) of the bytes of w (4 packed bytes) that are >= c, c < 128; exact for every byte value
__device__ __forceinline__ u32 ge_flags4(u32 w, u32 c4) {
  const u32 t = ((w & 0x7f7f7f7fu) | 0x80808080u) - c4;  // bit 7 of a byte survives iff its low 7 bits >= c
  return (t | w) & 0x80808080u;
}

template <bool ALL>
__global__ void __launch_bounds__(st::G::NT, ALL ? 3 : 4) k_stats_tile(StatsTileArgs a) {
  using namespace st;
  typedef st::Smem<ALL> Smem;
  using tile::H;
  using tile::PRE;
  constexpr u32 NT = G::NT, T = G::T, NSTAGE = G::NSTAGE;
  constexpr u32 HS = HBS;  // bins of the shared-memory histogram
  BSK_DYN_SMEM(Smem, smp);
  Smem &sm = *smp;
  const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const u32 n = a.n, n16 = n & ~15u;
  const bool fq = a.fastq != 0;

  for (u32 i = tid; i < HS; i += NT) sm.hist[i] = 0;
  if (tid == 0) {
    for (u32 s = 0; s < NSTAGE; s++) tma::mbar_init(&sm.full[s], 1);
    tma::fence_barrier_init();
    sm.n_rec = 0;
    sm.acc[0] = sm.acc[1] = sm.acc[2] = 0;
  }
  __syncthreads();
  if (tid == 0) {
    for (u32 p = 0; p < NSTAGE; p++) {
      const u32 tl = blockIdx.x + p * gridDim.x;
      if (tl < a.n_tiles) tile::issue_load<G>(a.in, n16, tl, sm.in[p], &sm.full[p]);
    }
  }
  u32 q20 = 0, q30 = 0, gaps = 0;  // per-thread partial counts, folded at the end
  const u32 c20 = (u32)(a.fq_offset + 20) * 0x01010101u, c30 = (u32)(a.fq_offset + 30) * 0x01010101u;

  u32 it = 0;
  for (u32 tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, it++) {
    const u32 s = it % NSTAGE, parity = (it / NSTAGE) & 1u;
    const u32 t0 = tile * T;
    u8 *stage = sm.in[s];
    const u8 *d = stage + PRE;
    const u32 lim = (n - t0 < T + H) ? n - t0 : T + H;
    const bool eof = (n - t0) <= T + H;
    {
      u32 g0, g1;
      if (tile::bulk_range<G>(tile, n16, g0, g1)) tma::mbar_wait(&sm.full[s], parity);
    }
    tile::fill_edges<G>(a.in, n, tile, stage);

    // one line of the list: does it open an owned record, how long is its sequence, is the record complete?
    struct Ev { bool own; u32 slen, k2, st; };  // st: 0 ok, 1 needs the rest of the halo, 2 outside the grammar
    auto evaluate = [&](u32 k, u32 n_lines, u32 slim) {
      Ev r{false, 0u, 0u, 0u};
      if (k > n_lines) return r;
      const u32 e0 = sm.ls[k];
      if (!(e0 & 0x8000u) || (e0 & 0x7fffu) >= T) return r;
      r.own = true;
      if (fq) {
        // "@h \n s \n +.. \n q \n" with |s| == |q| (SeqParser.Read on a 4-line record)
        if (k + 4 <= n_lines) {
          const u32 e1 = sm.ls[k + 1], e2 = sm.ls[k + 2], e3 = sm.ls[k + 3], e4 = sm.ls[k + 4];
          const u32 l1 = e1 & 0x7fffu, l2 = e2 & 0x7fffu, l3 = e3 & 0x7fffu, l4 = e4 & 0x7fffu;
          const u32 ql = l4 - 1 - l3;
          r.slen = l2 - 1 - l1;
          bool ok = ((e1 | e2 | e3) & 0x8000u) == 0;
          ok = ok && l3 > l2 + 1 && d[l2] == '+';  // the separator line starts with '+' (text after it is allowed)
          ok = ok && r.slen == ql && !(r.slen > 0 && d[l1] == '+');
          ok = ok && ((e4 & 0x8000u) || (eof && l4 >= lim));
          if (!ok) r.st = 2;
        } else {
          r.st = slim < lim ? 1u : 2u;
        }
      } else {
        // FASTA: every line up to the next record line (or the end of the file) is sequence
        u32 k2 = k + 1;
        while (k2 <= n_lines && !(sm.ls[k2] & 0x8000u)) k2++;
        if (k2 > n_lines && !(eof && slim == lim)) {
          r.st = slim < lim ? 1u : 2u;  // the record runs past the scanned part of the halo / past the halo
        } else {
          if (k2 > n_lines) k2 = n_lines;  // last record of the file: lines k+1 .. n_lines-1
          r.k2 = k2;
          r.slen = k2 > k + 1 ? ((sm.ls[k2] & 0x7fffu) - (sm.ls[k + 1] & 0x7fffu)) - (k2 - k - 1) : 0u;
        }
      }
      if (r.st == 0 && r.slen >= HB) r.st = 2;
      return r;
    };
    // complete owned records of a warp: lengths -> histogram.  Reads of one length all hit the same bin, which a
    // shared-memory atomic would serialise lane by lane: when every record of the warp has the length of the first
    // one, a single lane adds the count.  Every lane of the warp calls it (valid = this lane has a record).
    auto commit_hist = [&](bool valid, u32 slen) {
      const u32 bal = __ballot_sync(0xffffffffu, valid);
      if (!bal) return;
      const u32 leader = (u32)__ffs((int)bal) - 1u;
      const u32 first = __shfl_sync(0xffffffffu, slen, (int)leader);
      if (__all_sync(0xffffffffu, !valid || slen == first)) {
        if (lane == leader) {
          if (first < HS) atomicAdd(&sm.hist[first], (u32)__popc(bal));
          else atomicAdd((unsigned long long *)&a.hist[first], (unsigned long long)__popc(bal));
        }
      } else if (valid) {
        if (slen < HS) atomicAdd(&sm.hist[slen], 1u);
        else atomicAdd((unsigned long long *)&a.hist[slen], 1ull);
      }
    };
    // with -a the sequence / quality lines of a complete owned record become work items
    auto commit_items = [&](u32 k, const Ev &r) {
      if (!ALL || !r.slen) return;
      if (fq) {
        const u32 i0 = atomicAdd(&sm.n_item, 2u);
        if (i0 + 2 <= ICAP) {
          sm.item[i0] = (sm.ls[k + 1] & 0x7fffu) | (r.slen << 15);
          sm.item[i0 + 1] = (sm.ls[k + 3] & 0x7fffu) | (r.slen << 15) | (1u << 30);
        } else sm.bad = 1;
      } else {
        const u32 nl = r.k2 - k - 1;
        const u32 i0 = atomicAdd(&sm.n_item, nl);
        if (i0 + nl <= ICAP) {
          for (u32 j = 0; j < nl; j++) {
            const u32 b = sm.ls[k + 1 + j] & 0x7fffu, e = (sm.ls[k + 2 + j] & 0x7fffu) - 1u;
            sm.item[i0 + j] = b | ((e - b) << 15);
          }
        } else sm.bad = 1;
      }
    };

    u32 hs = a.scan_halo;
    bool bad = false;
    for (;;) {  // uniform
      const u32 slim = lim < T + hs ? lim : T + hs;
      const u32 n_lines = tile::scan_lines<G>(d, lim, slim, eof, t0, tile, fq, sm.ls, sm.wtot,
                                              [&]() { sm.bad = 0; sm.rescan = 0; sm.n_item = 0; });
      if (n_lines == 0xffffffffu) { bad = true; break; }
      // A short first-attempt scan may have to be repeated over the whole halo; nothing may be counted before that
      // is known.  With the whole halo scanned (direct) records are counted as they are evaluated.
      const bool direct = hs >= H;
      if (!direct && n_lines >= NT) {  // more lines than threads: no registers to park the evaluation in
        hs = H;
        __syncthreads();
        continue;
      }
      Ev mine{false, 0u, 0u, 0u};
      u32 n_new = 0;
      for (u32 kb = 0; kb <= n_lines; kb += NT) {  // one trip unless direct
        if (kb + warp * 32u > n_lines) continue;    // warp-uniform
        const Ev r = evaluate(kb + tid, n_lines, slim);
        const bool done = r.own && r.st == 0;
        if (r.own) {
          if (r.st == 1) sm.rescan = 1;
          else if (r.st == 2) sm.bad = 1;
        }
        if (direct) {
          commit_hist(done, r.slen);
          if (ALL && done) commit_items(kb + tid, r);
        } else {
          mine = r;
        }
        n_new += (u32)__popc(__ballot_sync(0xffffffffu, done));
      }
      __syncthreads();
      bad = sm.bad != 0 || (tile == 0 && !(sm.ls[0] & 0x8000u));
      if (bad) break;
      if (!direct && sm.rescan) {
        hs = H;
        __syncthreads();  // everybody has read the flags before thread 0 clears them again
        continue;
      }
      if (!direct) {  // (one trip above: n_lines < NT)
        const bool done = mine.own && mine.st == 0;
        commit_hist(done, mine.slen);
        if (ALL && done) commit_items(tid, mine);
      }
      if (lane == 0 && n_new) atomicAdd(&sm.n_rec, n_new);
      if (!direct) {
        __syncthreads();  // work items pushed by commit()
        bad = sm.bad != 0;
      }
      break;
    }
    if (bad) {
      if (tid == 0) atomicAdd((unsigned long long *)&a.st->counters[0], 1ull);
    } else if (ALL) {
      // ---- work items: 4 lanes per line (a 150-byte line is 10-11 chunks: three steps), one aligned 16-byte chunk per lane and step.  The bytes of the chunk that
      // belong to the line are selected by a flag mask (bit 7 of byte i set iff lo <= i < hi) that is ANDed with the
      // flags of the byte tests, so no byte is masked itself.
      // FASTQ items come in pairs (sequence line at the even index, quality line at the odd one): the sequence lines are
      // done first, then the quality lines, so that the two kinds of byte tests never share a warp.
      const u32 n_item = sm.n_item;
      const u32 g = tid >> 2, gl = tid & 3u;
      const u32 npass = fq ? 2u : 1u;
      for (u32 pass = 0; pass < npass; pass++)
      for (u32 ib = fq ? 2u * g + pass : g; ib < n_item; ib += fq ? NT / 2 : NT / 4) {
        const u32 e = sm.item[ib];
        const u32 p = e & 0x7fffu, L = (e >> 15) & 0x7fffu, kind = e >> 30;
        const u32 c1 = (p + L + 15u) & ~15u;
        for (u32 c = (p & ~15u) + 16u * gl; c < c1; c += 64u) {
          const uint4 v = *reinterpret_cast<const uint4 *>(d + c);
          const u32 lo = p > c ? p - c : 0u, hi = p + L - c < 16u ? p + L - c : 16u;
          const u32 lo4 = lo * 0x01010101u, hi4 = hi * 0x01010101u;
          const u32 w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (u32 k2 = 0; k2 < 4; k2++) {
            const u32 w = w4[k2];
            const u32 idx = (0x03020100u + k2 * 0x04040404u) | 0x80808080u;  // byte indices of the word in the chunk
            const u32 m7 = (idx - lo4) & ~(idx - hi4) & 0x80808080u;
            if (kind) {
              q20 += (u32)__popc(ge_flags4(w, c20) & m7);
              q30 += (u32)__popc(ge_flags4(w, c30) & m7);
            } else if (!a.gap_below_40 || ((((w | (w >> 1)) & 0x40404040u) | ((m7 >> 1) ^ 0x40404040u)) != 0x40404040u)) {
              // only words holding a byte below 0x40 INSIDE the line can hold one of the (sub-'@') gap letters; the
              // bytes around the line (its '\n' is below 0x40) must not count, or some lane of every warp comes here
              u32 f = 0;
#pragma unroll
              for (int j = 0; j < 4; j++)
                if (j < a.n_gap) {
                  const u32 x = w ^ a.gap[j];
                  f |= ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;  // distinct letters: disjoint flags
                }
              gaps += (u32)__popc(f & m7);
            }
          }
        }
      }
    }
    __syncthreads();  // every thread is done with the stage (and with ls / item) before it is refilled
    if (tid == 0) {
      const u32 tn = tile + NSTAGE * gridDim.x;
      if (tn < a.n_tiles) tile::issue_load<G>(a.in, n16, tn, sm.in[s], &sm.full[s]);
    }
  }
  // ---- fold the CTA's results into the global ones
  if (ALL) {
    for (int off = 16; off > 0; off >>= 1) {
      q20 += __shfl_xor_sync(0xffffffffu, q20, off);
      q30 += __shfl_xor_sync(0xffffffffu, q30, off);
      gaps += __shfl_xor_sync(0xffffffffu, gaps, off);
    }
    if (lane == 0) {
      atomicAdd(&sm.acc[0], (unsigned long long)q20);
      atomicAdd(&sm.acc[1], (unsigned long long)q30);
      atomicAdd(&sm.acc[2], (unsigned long long)gaps);
    }
  }
  __syncthreads();
  for (u32 i = tid; i < HS; i += NT) {
    const u32 c = sm.hist[i];
    if (c) atomicAdd((unsigned long long *)&a.hist[i], (unsigned long long)c);
  }
  if (tid == 0) {
    atomicAdd((unsigned long long *)&a.st->counters[5], (unsigned long long)sm.n_rec);
    if (ALL) {
      atomicAdd((unsigned long long *)&a.st->counters[1], sm.acc[0]);
      atomicAdd((unsigned long long *)&a.st->counters[2], sm.acc[1]);
      atomicAdd((unsigned long long *)&a.st->counters[3], sm.acc[2]);
    }
  }
}

u32 stats_tile_bins() { return st::HB; }

void stats_tile(const u8 *in, u32 n, u64 *hist, DevStatus *st, int fastq, int all, int fq_offset, const u8 *gap_letters,
                int n_gap, u32 scan_halo, int n_sm, cudaStream_t s) {
  StatsTileArgs a;
  a.in = in;
  a.n = n;
  a.hist = hist;
  a.st = st;
  a.n_tiles = (n + st::G::T - 1) / st::G::T;
  a.fastq = fastq;
  a.all = all;
  a.fq_offset = fq_offset;
  for (int i = 0; i < 4; i++) a.gap[i] = i < n_gap ? (u32)gap_letters[i] * 0x01010101u : 0u;
  a.n_gap = n_gap;
  a.gap_below_40 = 1;
  for (int i = 0; i < n_gap; i++)
    if (gap_letters[i] >= 0x40) a.gap_below_40 = 0;
  scan_halo = (scan_halo + 15u) & ~15u;
  a.scan_halo = scan_halo < 256u ? 256u : (scan_halo > tile::H ? tile::H : scan_halo);
  const int ctas = all ? 3 : 4;
  u32 grid = (u32)n_sm * ctas;
  if (grid > a.n_tiles) grid = a.n_tiles;
  if (grid == 0) return;
  const size_t smem = (all ? sizeof(st::Smem<true>) : sizeof(st::Smem<false>)) + 16;
#ifndef BSK_EMU
  // the opt-in to > 48 KiB of dynamic shared memory is per device (a process may hold ctxs on several GPUs)
  static bool attr_set[64][2] = {{false, false}};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev][all ? 1 : 0]) {
    if (all) cudaFuncSetAttribute(k_stats_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else cudaFuncSetAttribute(k_stats_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set[dev][all ? 1 : 0] = true;
  }
#endif
  if (all) BSK_LAUNCH(k_stats_tile<true>, grid, st::G::NT, smem, s, a);
  else BSK_LAUNCH(k_stats_tile<false>, grid, st::G::NT, smem, s, a);
}

}  // namespace k
}  // namespace bsk
