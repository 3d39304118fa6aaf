// k_rmdup_tile.cu -- front half of `rmdup` on 4-line FASTQ reads in ONE streaming pass: record index + parse +
// subject hashing, i.e. what k_index_count / k_index_fill / k_parse_records / k_rmdup_hash do in four passes.
//
//   PlainFile + ReadFixer + SeqParser.Read   bigseqkit/helper.go:148-178, bigseqkit-lib/helper.go:41-66, 219-325
//   RmDupPrepare.Call                        bigseqkit-lib/rmdup.go:43-90   key = int64(xxhash.Sum64(subject)),
//                                            subject = seq (-s) | Name (-n) | ID (default)
//
// Same skeleton as k_stats_tile.cu (tile_common.cuh).  Every owned record is hashed from shared memory by four
// lanes (XXH64 seed 0 = the reference's key, and FP64 for the 128-bit fingerprint, xxh64.cuh; 8-byte little-endian
// words assembled from aligned 32-bit loads) and leaves a 32-byte slot {key, fingerprint, record start, head / id / seq lengths};
// k_rmdup_tile_compact turns the per-tile slot lists into the dense per-record arrays (keys, fingerprints and the
// RecArrays of the general path), after which table insert / resolve / k_emit_contig run unchanged.
// Records outside the grammar ("+name" lines, multi-line, longer than the halo, --ignore-case) -> general path.
#include "kernels.h"
#include "tile_common.cuh"
#include "xxh64.cuh"

namespace bsk {
namespace k {

namespace rt {
typedef tile::Geo<512, 3, 3, 2, 3072> G;  // (4 CTAs / SM at 32 registers measured no faster: 0.855 vs 0.823 ms per GiB)
constexpr u32 RCAP = 192;  // record slots per tile (20 KiB tile: records of >= 107 bytes on average)

struct Smem {
  u8 in[G::NSTAGE][G::STAGE];
  u64 full[G::NSTAGE];
  u16 ls[G::LCAP + 8];
  u16 r_line[RCAP];
  u32 wtot[G::NWARP];
  u32 wcnt[G::NWARP];
  u32 bad;
};
}  // namespace rt

struct RmdupSlot {
  u64 key, fp;
  u32 rec_start;       // global offset of the '@'
  u16 hl, idl, sl, pad;
};
static_assert(sizeof(RmdupSlot) == 32, "slot layout");

struct RmdupTileArgs {
  const u8 *in;
  u32 n;
  RmdupSlot *slots;    // [n_tiles * RCAP]
  u32 *tile_cnt;       // [n_tiles + 1]
  DevStatus *st;       // counters[0] = declined tiles
  u32 n_tiles;
  int subject;         // 0 sequence, 1 name, 2 id; -1: index + parse only (no hashing)
};

// 8 / 4 little-endian bytes at region offset off + i from aligned 32-bit words (reads up to 3 bytes of slack)
struct GetWords {
  const u8 *base;  // 4-byte aligned
  u32 off;
  __host__ __device__ __forceinline__ u8 operator()(u32 i) const { return base[off + i]; }
};
}  // namespace k

template <>
__host__ __device__ __forceinline__ u64 xx_rd64<k::GetWords>(const k::GetWords &g, u32 i) {
  const u32 a = g.off + i;
  const u32 *w = reinterpret_cast<const u32 *>(g.base + (a & ~3u));
  const u32 sh = (a & 3u) * 8u;
  const u32 w0 = w[0], w1 = w[1], w2 = sh ? w[2] : 0u;
  const u64 lo = ((u64)w1 << 32 | w0) >> sh, hi = ((u64)w2 << 32 | w1) >> sh;  // 64-bit shifts: usable on host and device
  return (lo & 0xffffffffull) | (hi << 32);
}
template <>
__host__ __device__ __forceinline__ u64 xx_rd32<k::GetWords>(const k::GetWords &g, u32 i) {
  const u32 a = g.off + i;
  const u32 *w = reinterpret_cast<const u32 *>(g.base + (a & ~3u));
  const u32 sh = (a & 3u) * 8u;
  const u32 w0 = w[0], w1 = sh ? w[1] : 0u;
  return (((u64)w1 << 32 | w0) >> sh) & 0xffffffffull;
}

namespace k {

// parseHeadIDAndDesc, default regexp (bigseqkit-lib/helper.go:329-369): ID = head up to the first ' ' at index > 0,
// else up to the first '\t' at index > 0, else the whole head
__device__ __forceinline__ u32 rt_id_len(const u8 *h, u32 e) {
  u32 i = e;
  for (u32 t = 0; t < e; t++)
    if (h[t] == ' ') { i = t; break; }
  if (i == e || i == 0) {
    i = e;
    for (u32 t = 0; t < e; t++)
      if (h[t] == '\t') { i = t; break; }
    if (i == 0) i = e;
  }
  return i;
}

__global__ void __launch_bounds__(rt::G::NT, rt::G::CTAS) k_rmdup_tile(RmdupTileArgs a) {
  using namespace rt;
  using tile::H;
  using tile::PRE;
  constexpr u32 NT = G::NT, T = G::T, NSTAGE = G::NSTAGE, NWARP = G::NWARP;
  BSK_DYN_SMEM(Smem, smp);
  Smem &sm = *smp;
  const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const u32 n = a.n, n16 = n & ~15u;

  if (tid == 0) {
    for (u32 s = 0; s < NSTAGE; s++) tma::mbar_init(&sm.full[s], 1);
    tma::fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (u32 p = 0; p < NSTAGE; p++) {
      const u32 tl = blockIdx.x + p * gridDim.x;
      if (tl < a.n_tiles) tile::issue_load<G>(a.in, n16, tl, sm.in[p], &sm.full[p]);
    }
  }
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
    const u32 n_lines = tile::scan_lines<G>(d, lim, lim, eof, t0, tile, true, sm.ls, sm.wtot, [&]() { sm.bad = 0; });
    bool bad = n_lines == 0xffffffffu;
    u32 n_own = 0;
    if (!bad) {
      // ---- owned records in input order: "@h \n s \n + \n q \n" with |s| == |q| (bare '+': the record is printed as is)
      u32 run = 0;
      for (u32 kb = 0; kb <= n_lines; kb += NT) {  // uniform trip count
        const u32 k = kb + tid;
        bool own = false;
        if (k <= n_lines) {
          const u32 e0 = sm.ls[k];
          if ((e0 & 0x8000u) && (e0 & 0x7fffu) < T) {
            own = true;
            bool ok = k + 4 <= n_lines;
            if (ok) {
              const u32 e1 = sm.ls[k + 1], e2 = sm.ls[k + 2], e3 = sm.ls[k + 3], e4 = sm.ls[k + 4];
              const u32 l1 = e1 & 0x7fffu, l2 = e2 & 0x7fffu, l3 = e3 & 0x7fffu, l4 = e4 & 0x7fffu;
              const u32 sl = l2 - 1 - l1, ql = l4 - 1 - l3;
              ok = ((e1 | e2 | e3) & 0x8000u) == 0;
              ok = ok && (l3 - l2 == 2) && d[l2] == '+';
              ok = ok && sl == ql && !(sl > 0 && d[l1] == '+');
              ok = ok && ((e4 & 0x8000u) || (eof && l4 >= lim));
              ok = ok && !(eof && l4 > lim);  // a last record without its final newline is not printed "as is"
            }
            if (!ok) sm.bad = 1;
          }
        }
        const u32 bal = __ballot_sync(0xffffffffu, own);
        if (lane == 0) sm.wcnt[warp] = (u32)__popc(bal);
        __syncthreads();
        u32 wb = 0, tot = 0;
        for (u32 w = 0; w < NWARP; w++) {
          const u32 c = sm.wcnt[w];
          if (w < warp) wb += c;
          tot += c;
        }
        if (own) {
          const u32 r = run + wb + (u32)__popc(bal & ((1u << lane) - 1u));
          if (r < RCAP) sm.r_line[r] = (u16)k;
        }
        run += tot;
        __syncthreads();
      }
      n_own = run;
      bad = sm.bad != 0 || n_own > RCAP || (tile == 0 && !(sm.ls[0] & 0x8000u));
    }
    if (bad) {
      if (tid == 0) atomicAdd((unsigned long long *)&a.st->counters[0], 1ull);
    } else {
      // ---- four lanes per record: lane i takes word i of every 32-byte stripe of the subject (XXH64's four
      // accumulators are independent, FP64's lanes likewise); the tail and the final mix are done by all four alike
      const u32 gl = tid & 3u, g0lane = lane & ~3u;
      for (u32 rb = 0; rb < n_own; rb += NT / 4) {  // uniform trip count
        const u32 r = rb + (tid >> 2);
        const bool act = r < n_own;
        u32 p0 = 0, hl = 0, sl = 0, idl = 0, so = 0, slen = 0;
        if (act) {
          const u32 k = sm.r_line[r];
          p0 = sm.ls[k] & 0x7fffu;
          const u32 l1 = sm.ls[k + 1] & 0x7fffu, l2 = sm.ls[k + 2] & 0x7fffu;
          hl = l1 - 1 - (p0 + 1);
          sl = l2 - 1 - l1;
          so = l1;
          slen = sl;
          if (a.subject == 1) { so = p0 + 1; slen = hl; }
          else if (a.subject == 2) { idl = rt_id_len(d + p0 + 1, hl); so = p0 + 1; slen = idl; }
          if (a.subject < 0) slen = 0;
        }
        const GetWords gw{d, so};
        u64 v = gl == 0 ? XXP1 + XXP2 : (gl == 1 ? XXP2 : (gl == 2 ? 0ull : 0ull - XXP1));
        u64 g = fp_lane_init(gl);
        const u32 ns = slen >> 5;
        for (u32 s2 = 0; s2 < ns; s2++) {
          const u64 w = xx_rd64(gw, 32u * s2 + 8u * gl);
          v = xx_round(v, w);
          g = fp_lane_step(g, w);
        }
        const u64 v1 = __shfl_sync(0xffffffffu, v, g0lane), v2 = __shfl_sync(0xffffffffu, v, g0lane + 1u);
        const u64 v3 = __shfl_sync(0xffffffffu, v, g0lane + 2u), v4 = __shfl_sync(0xffffffffu, v, g0lane + 3u);
        const u64 f0 = __shfl_sync(0xffffffffu, g, g0lane), f1 = __shfl_sync(0xffffffffu, g, g0lane + 1u);
        const u64 f2 = __shfl_sync(0xffffffffu, g, g0lane + 2u), f3 = __shfl_sync(0xffffffffu, g, g0lane + 3u);
        u64 h;
        if (ns) {
          h = xx_rotl(v1, 1) + xx_rotl(v2, 7) + xx_rotl(v3, 12) + xx_rotl(v4, 18);
          h = xx_merge(h, v1);
          h = xx_merge(h, v2);
          h = xx_merge(h, v3);
          h = xx_merge(h, v4);
        } else {
          h = XXP5;
        }
        h += (u64)slen;
        u64 t = XXP5;
        key_fp_tail(gw, 32u * ns, slen, h, t);
        if (act && gl == 0) {
          RmdupSlot sl_;
          sl_.key = a.subject >= 0 ? xx_avalanche(h) : 0;
          sl_.fp = a.subject >= 0 ? fp_finish(f0, f1, f2, f3, t, slen) : 0;
          sl_.rec_start = t0 + p0;
          sl_.hl = (u16)hl;
          sl_.idl = (u16)idl;
          sl_.sl = (u16)sl;
          sl_.pad = 0;
          a.slots[(size_t)tile * RCAP + r] = sl_;
        }
      }
      if (tid == 0) a.tile_cnt[tile] = n_own;
    }
    // Every thread is done with the stage before it is refilled -- but only the refilling warp waits for that: the
    // others arrive and go on to scan the next tile (other stage) while the hashing lanes finish; nothing of this
    // tile is overwritten before the next full barrier (inside scan_lines), which every thread joins.
    if (warp != 0) {
      tma::named_arrive(1, NT);
    } else {
      tma::named_sync(1, NT);
      if (tid == 0) {
        const u32 tn = tile + NSTAGE * gridDim.x;
        if (tn < a.n_tiles) tile::issue_load<G>(a.in, n16, tn, sm.in[s], &sm.full[s]);
      }
    }
  }
}

// dense per-record arrays from the slot lists: one warp per tile
__global__ void k_rmdup_tile_compact(const RmdupSlot *__restrict__ slots, const u32 *__restrict__ tile_cnt,
                                     const u64 *__restrict__ tile_base, u32 n_tiles, u64 *__restrict__ keys,
                                     u64 *__restrict__ fps, RecArrays ra, u32 *__restrict__ id_len) {
  const u32 tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (tile >= n_tiles) return;
  const u32 c = tile_cnt[tile];
  const u64 b = tile_base[tile];
  for (u32 r = lane; r < c; r += 32) {
    const RmdupSlot s = slots[(size_t)tile * rt::RCAP + r];
    const u64 i = b + r;
    if (keys) {
      keys[i] = s.key;
      fps[i] = s.fp;
    }
    const u32 ho = s.rec_start + 1, so = ho + s.hl + 1;
    ra.head_off[i] = ho;
    ra.head_len[i] = s.hl;
    ra.seq_off[i] = so;
    ra.seq_len[i] = s.sl;
    ra.qual_off[i] = so + s.sl + 3;
    ra.qual_len[i] = s.sl;
    ra.seq_line0[i] = ra.seq_line1[i] = ra.qual_line0[i] = ra.qual_line1[i] = 0;  // no line index on this path
    if (id_len) id_len[i] = s.idl;
  }
}

u32 rmdup_tile_tiles(u32 n) { return (n + rt::G::T - 1) / rt::G::T; }
u32 rmdup_tile_slot_stride() { return rt::RCAP; }
size_t rmdup_tile_slot_bytes() { return sizeof(RmdupSlot); }

void rmdup_tile(const u8 *in, u32 n, void *slots, u32 *tile_cnt, DevStatus *st, int subject, int n_sm, cudaStream_t s) {
  RmdupTileArgs a;
  a.in = in;
  a.n = n;
  a.slots = static_cast<RmdupSlot *>(slots);
  a.tile_cnt = tile_cnt;
  a.st = st;
  a.n_tiles = rmdup_tile_tiles(n);
  a.subject = subject;
  const size_t smem = sizeof(rt::Smem) + 16;
#ifndef BSK_EMU
  // the opt-in to > 48 KiB of dynamic shared memory is per device (a process may hold ctxs on several GPUs)
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaFuncSetAttribute(k_rmdup_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set[dev] = true;
  }
#endif
  u32 grid = (u32)n_sm * rt::G::CTAS;
  if (grid > a.n_tiles) grid = a.n_tiles;
  if (grid == 0) return;
  BSK_LAUNCH(k_rmdup_tile, grid, rt::G::NT, smem, s, a);
}

void rmdup_tile_compact(const void *slots, const u32 *tile_cnt, const u64 *tile_base, u32 n_tiles, u64 *keys, u64 *fps,
                        RecArrays ra, u32 *id_len, cudaStream_t s) {
  if (!n_tiles) return;
  const u64 threads = (u64)n_tiles * 32;
  BSK_LAUNCH_FLAT(k_rmdup_tile_compact, (u32)((threads + 255) / 256), 256, 0, s, static_cast<const RmdupSlot *>(slots), tile_cnt,
                  tile_base, n_tiles, keys, fps, ra, id_len);
}

}  // namespace k
}  // namespace bsk
