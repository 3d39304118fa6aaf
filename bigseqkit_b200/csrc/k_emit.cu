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

// last record r in [0, n_rec) with off[r] <= o, found by a whole warp: 32 probes per step instead of one
__device__ __forceinline__ u32 warp_search(const u64 *__restrict__ off, u32 n_rec, u64 o) {
  const u32 lane = threadIdx.x & 31;
  u32 lo = 0, hi = n_rec;  // off[lo] <= o ; hi == n_rec or off[hi] > o
  while (hi - lo > 1) {
    const u32 span = hi - lo;
    const u32 step = (span + 32) / 33;
    const u32 idx = lo + (lane + 1) * step;
    const bool le = idx < hi && off[idx] <= o;
    const u32 cnt = (u32)__popc(__ballot_sync(0xffffffffu, le));  // off[] is monotone: the true lanes are a prefix
    const u32 nlo = lo + cnt * step;
    const u32 nhi = lo + (cnt + 1) * step;
    lo = nlo;
    if (nhi < hi) hi = nhi;
  }
  return lo;
}

static const u32 kEmitChunks = 4;  // 16-byte chunks per thread: a CTA of 256 threads writes 16 KiB

// record range [r0, r1] that covers the CTA's output bytes, found once per CTA
__device__ __forceinline__ void cta_record_range(const u64 *__restrict__ off, u32 n_rec, u64 o0, u64 total, u32 *s_r,
                                                 u64 cta_bytes = 256ull * 16ull * kEmitChunks) {
  const u32 warp = threadIdx.x >> 5;
  if (warp < 2) {
    u64 o = warp == 0 ? o0 : o0 + cta_bytes - 1;
    if (o >= total) o = total - 1;
    const u32 r = warp_search(off, n_rec, o);
    if ((threadIdx.x & 31) == 0) s_r[warp] = r;
  }
  __syncthreads();
}

__device__ __forceinline__ u32 search_in(const u64 *__restrict__ off, u32 lo, u32 hi, u64 o) {  // off[lo] <= o < off[hi]
  while (hi - lo > 1) {
    const u32 mid = lo + ((hi - lo) >> 1);
    if (off[mid] <= o) lo = mid;
    else hi = mid;
  }
  return lo;
}

// 16-byte window of base[src .. src+16) at any alignment: five aligned word loads + funnel shifts.  Words that
// start at or beyond `limit` are not touched (read as 0).
__device__ __forceinline__ void window16(const u8 *__restrict__ base, u64 src, u64 limit, u32 ww[4]) {
  const u64 a0 = src & ~3ull;
  const u32 *w = reinterpret_cast<const u32 *>(base + a0);
  const u32 sh = (u32)(src & 3ull) * 8u;
  u32 a, b, cc, d, e;
  if (a0 + 20 <= limit) {
    a = w[0]; b = w[1]; cc = w[2]; d = w[3]; e = w[4];
  } else {
    a = a0 < limit ? w[0] : 0u;
    b = a0 + 4 < limit ? w[1] : 0u;
    cc = a0 + 8 < limit ? w[2] : 0u;
    d = a0 + 12 < limit ? w[3] : 0u;
    e = a0 + 16 < limit ? w[4] : 0u;
  }
  ww[0] = __funnelshift_r(a, b, sh);
  ww[1] = __funnelshift_r(b, cc, sh);
  ww[2] = __funnelshift_r(cc, d, sh);
  ww[3] = __funnelshift_r(d, e, sh);
}

// OR bytes [shift, shift + cnt) of the window ww into the chunk words w
__device__ __forceinline__ void merge16(u32 w[4], const u32 ww[4], u32 shift, u32 cnt) {
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int l = (int)shift - 4 * q, h = (int)(shift + cnt) - 4 * q;  // bytes [l, h) of word q are taken
    if (h <= 0 || l >= 4) continue;
    u32 m = 0xffffffffu;
    if (l > 0) m &= 0xffffffffu << (8 * l);
    if (h < 4) m &= 0xffffffffu >> (8 * (4 - h));
    w[q] |= ww[q] & m;
  }
}

// The piece of record output that starts at position p of the record: a literal byte, a run copied straight from a
// source buffer, or a run that needs the byte path (reversed and / or mapped sequence).
struct Piece {
  u32 len;        // bytes in the piece from p on (>= 1)
  int kind;       // 0 literal, 1 run of source bytes
  u8 lit;
  const u8 *base; // kind 1: source buffer and offset of the byte that comes out FIRST
  u64 src;
  bool rev;       // kind 1: the following output bytes come from DEcreasing source offsets
  bool map;       // kind 1: bytes go through the sequence byte map
};

__device__ __forceinline__ Piece piece_at(const RecViews &v, const EmitCfg &c, const RecOut &o, u32 p, bool has_lut) {
  Piece pc;
  pc.base = nullptr;
  pc.src = 0;
  pc.rev = false;
  pc.map = false;
  if (p < o.np) {
    u32 q = p;
    if (c.marker) {
      if (q == 0) { pc.kind = 0; pc.len = 1; pc.lit = c.marker; return pc; }
      q--;
    }
    if (q < o.name_len) { pc.kind = 1; pc.len = o.name_len - q; pc.base = v.in; pc.src = (u64)o.name_off + q; return pc; }
    pc.kind = 0; pc.len = 1; pc.lit = '\n';
    return pc;
  }
  p -= o.np;
  if (p < o.ns) {
    if (p == o.wl) { pc.kind = 0; pc.len = 1; pc.lit = '\n'; return pc; }
    u32 j = p, run = o.seq_len - p;
    if (c.width > 0) {
      const u32 line = p / (c.width + 1u), col = p - line * (c.width + 1u);
      if (col == c.width) { pc.kind = 0; pc.len = 1; pc.lit = '\n'; return pc; }
      j = line * c.width + col;
      run = c.width - col;
      if (run > o.seq_len - j) run = o.seq_len - j;
    }
    pc.len = run;
    pc.kind = 1;
    pc.base = v.seqb;
    pc.map = has_lut;
    pc.rev = c.reverse != 0;
    pc.src = c.reverse ? (u64)o.seq_off + (o.seq_len - 1u - j) : (u64)o.seq_off + j;
    return pc;
  }
  p -= o.ns;
  if (c.plus_line) {
    if (p == 0) { pc.kind = 0; pc.len = 1; pc.lit = '+'; return pc; }
    if (p == 1) { pc.kind = 0; pc.len = 1; pc.lit = '\n'; return pc; }
    p -= 2;
  }
  if (p < o.qual_len) {
    pc.len = o.qual_len - p;
    pc.kind = 1;
    pc.base = v.qualb;
    pc.rev = c.reverse != 0;
    pc.src = c.reverse ? (u64)o.qual_off + (o.qual_len - 1u - p) : (u64)o.qual_off + p;
    return pc;
  }
  pc.kind = 0; pc.len = 1; pc.lit = '\n';
  return pc;
}

// ---- the CTA's slice of the record table in shared memory + a chunk -> record map (k_emit, k_emit_contig)
// s_off[i] = off[r0 + i] - o0 for the records r0 .. r0 + nr (clamped to int); s_map[c] = last record WITH output that
// starts at or before 16-byte chunk c of the CTA (every such record marks the first chunk that starts inside it, a
// prefix maximum fills the rest; chunks in front of the first mark get 0 and walk forward from there).
static const u32 kSliceCap = 1024;
template <u32 CH>
__device__ __forceinline__ void cta_slice_map(const u64 *__restrict__ off, u32 r0, u32 nr, u64 o0, int *s_off, u32 *s_map,
                                              u32 *s_wmax) {
  constexpr u32 NCHUNK = 256 * CH;
  const u32 tid = threadIdx.x;
  for (u32 i = tid; i <= nr; i += 256) {
    const long long rel = (long long)off[r0 + i] - (long long)o0;
    s_off[i] = rel > 0x7fffffffll ? 0x7fffffff : (rel < -0x7fffffffll ? -0x7fffffff : (int)rel);
  }
#pragma unroll
  for (u32 q = 0; q < CH; q++) s_map[CH * tid + q] = 0;
  __syncthreads();
  for (u32 i = tid; i < nr; i += 256) {
    const int b = s_off[i];
    if (b > 0 && s_off[i + 1] > b) {
      const u32 c = ((u32)b + 15u) >> 4;
      if (c < NCHUNK) atomicMax(&s_map[c], i);
    }
  }
  __syncthreads();
  // inclusive prefix maximum: CH consecutive entries per thread, warp scan, warp totals
  u32 mx[CH];
  u32 run = 0;
#pragma unroll
  for (u32 q = 0; q < CH; q++) {
    const u32 x = s_map[CH * tid + q];
    run = x > run ? x : run;
    mx[q] = run;
  }
  u32 inc = run;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u32 y = __shfl_up_sync(0xffffffffu, inc, d);
    if ((int)(tid & 31u) >= d && y > inc) inc = y;
  }
  if ((tid & 31u) == 31u) s_wmax[tid >> 5] = inc;
  u32 before = __shfl_up_sync(0xffffffffu, inc, 1);
  if ((tid & 31u) == 0u) before = 0;
  __syncthreads();
  for (u32 w = 0; w < (tid >> 5); w++) before = s_wmax[w] > before ? s_wmax[w] : before;
#pragma unroll
  for (u32 q = 0; q < CH; q++) s_map[CH * tid + q] = mx[q] > before ? mx[q] : before;
  __syncthreads();
}

// 16 output bytes from position p of record ro on: every run of source bytes (names, sequence / quality lines in either
// direction) moves as a 16-byte window -- byte-reversed with four PRMT for --reverse, mapped through the 256-byte table
// for --complement / case / dna2rna; only the literal bytes are placed one by one.  next(r) steps to the following
// record with output.
template <class Next>
__device__ __forceinline__ void emit_chunk_walk(const RecViews &v, const EmitCfg &c, RecOut ro, u32 p, long long rend, long long pos,
                                                long long oend, const u8 *__restrict__ lut, u64 in_limit, u64 seq_limit,
                                                u64 qual_limit, u32 w[4], Next next) {
  const bool has_lut = lut != nullptr;
  const long long o = pos;
  while (pos < oend) {
    if (pos >= rend) {
      next(ro, rend, pos);
      p = 0;
    }
    const Piece pc = piece_at(v, c, ro, p, has_lut);
    u32 cnt = pc.len;
    if (cnt > oend - pos) cnt = (u32)(oend - pos);
    const u32 shift = (u32)(pos - o);
    if (pc.kind == 0) {
      w[shift >> 2] |= (u32)pc.lit << (8 * (shift & 3));
    } else {
      // Window of 16 source bytes laid out so that chunk byte shift+t holds the t-th byte of the run: forward runs
      // start the window at src - shift; reversed runs end it at src + shift and are byte-reversed after the load.
      const u64 limit = pc.base == v.in ? in_limit : (pc.base == v.seqb ? seq_limit : qual_limit);
      const bool fits = pc.rev ? (pc.src + shift >= 15) : (pc.src >= shift);
      if (fits) {
        u32 ww[4];
        if (!pc.rev) {
          window16(pc.base, pc.src - shift, limit, ww);
        } else {
          u32 t4[4];
          window16(pc.base, pc.src + shift - 15, limit, t4);
          ww[0] = __byte_perm(t4[3], 0, 0x0123);
          ww[1] = __byte_perm(t4[2], 0, 0x0123);
          ww[2] = __byte_perm(t4[1], 0, 0x0123);
          ww[3] = __byte_perm(t4[0], 0, 0x0123);
        }
        if (pc.map) {
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const u32 x = ww[q];
            ww[q] = (u32)lut[x & 0xffu] | ((u32)lut[(x >> 8) & 0xffu] << 8) | ((u32)lut[(x >> 16) & 0xffu] << 16) |
                    ((u32)lut[x >> 24] << 24);
          }
        }
        merge16(w, ww, shift, cnt);
      } else {
        for (u32 t = 0; t < cnt; t++) w[(shift + t) >> 2] |= (u32)rec_byte(v, c, ro, p + t, lut) << (8 * ((shift + t) & 3));
      }
    }
    pos += cnt;
    p += cnt;
  }
}

// One aligned 16-byte store per thread and step.  The CTA's slice of the record table is staged in shared memory with a
// chunk -> record map (no binary search per chunk).  A chunk that lies inside ONE run of source bytes (the middle of a
// name, sequence or quality line: most of them) is one window and one store; the chunks that cross a piece or record
// boundary -- one or two lanes of every warp, which would make the whole warp run the general walk -- go on a list and
// are done side by side afterwards.  Slices of more than kSliceCap records (tiny records) take the walk for every chunk.
template <int MB>
__global__ void __launch_bounds__(256, MB) k_emit(RecViews v, EmitCfg c, const u64 *__restrict__ off, u8 *__restrict__ out,
                                              u64 total, const u8 *__restrict__ lut, u64 in_limit, u64 seq_limit, u64 qual_limit) {
  constexpr u32 NCHUNK = 256 * kEmitChunks;
  __shared__ u32 s_r[2];
  __shared__ int s_off[kSliceCap + 2];
  __shared__ u32 s_map[NCHUNK];
  __shared__ u32 s_wmax[8];
  __shared__ unsigned short s_slow[NCHUNK];
  __shared__ u32 s_nslow;
  constexpr u32 kRecCap = 256;
  __shared__ u32 s_rec[kRecCap][6];
  const u32 tid = threadIdx.x;
  if (tid == 0) s_nslow = 0;
  const u64 o0 = (u64)blockIdx.x * NCHUNK * 16ull;
  cta_record_range(off, v.n_rec, o0, total, s_r);
  const u32 r0 = s_r[0], nr = s_r[1] - s_r[0] + 1;
  const bool has_lut = lut != nullptr;
  if (nr <= kSliceCap) {
    // the views of the slice's records beside (up to kRecCap of them: reads and longer records; else from global memory)
    const bool recs_staged = nr <= kRecCap;
    if (recs_staged) {
      for (u32 i = tid; i < nr; i += 256) {
        const u32 r = r0 + i;
        s_rec[i][0] = v.name_off[r];
        s_rec[i][1] = v.name_len[r];
        s_rec[i][2] = v.seq_off[r];
        s_rec[i][3] = v.seq_len[r];
        s_rec[i][4] = c.print_qual ? v.qual_off[r] : 0u;
        s_rec[i][5] = c.print_qual ? v.qual_len[r] : 0u;
      }
    }
    auto rec_of = [&](u32 i) {
      if (!recs_staged) return load_rec(v, c, r0 + i);
      RecOut o;
      o.name_off = s_rec[i][0];
      o.name_len = s_rec[i][1];
      o.seq_off = s_rec[i][2];
      o.seq_len = s_rec[i][3];
      o.qual_off = s_rec[i][4];
      o.qual_len = s_rec[i][5];
      o.np = c.print_name ? (c.marker ? 1u : 0u) + o.name_len + 1u : 0u;
      o.wl = wrap_len(o.seq_len, c.width);
      o.ns = c.print_seq ? o.wl + 1u : 0u;
      o.nq = c.print_qual ? (c.plus_line ? 2u : 0u) + o.qual_len + 1u : 0u;
      return o;
    };
    cta_slice_map<kEmitChunks>(off, r0, nr, o0, s_off, s_map, s_wmax);
    auto locate = [&](u32 cidx, u32 &i, int &rbeg, int &rend) {
      const int ro = (int)(cidx * 16u);
      i = s_map[cidx];
      rbeg = s_off[i];
      rend = s_off[i + 1];
      while (ro >= rend) {  // entry 0 of the map may be a record without output
        i++;
        rbeg = rend;
        rend = s_off[i + 1];
      }
    };
#pragma unroll 1
    for (u32 ch = 0; ch < kEmitChunks; ch++) {
      const u32 cidx = ch * 256u + tid;
      const u64 o = o0 + (u64)cidx * 16ull;
      if (o >= total) break;
      bool done = false;
      if (o + 16 <= total) {
        u32 i;
        int rbeg, rend;
        locate(cidx, i, rbeg, rend);
        const int ro = (int)(cidx * 16u);
        if (rend >= ro + 16) {
          const RecOut rc = rec_of(i);
          const Piece pc = piece_at(v, c, rc, (u32)(ro - rbeg), has_lut);
          const u64 limit = pc.base == v.in ? in_limit : (pc.base == v.seqb ? seq_limit : qual_limit);
          if (pc.kind == 1 && pc.len >= 16u && (!pc.rev || pc.src >= 15u)) {  // the whole chunk is one run of source bytes
            u32 w[4];
            if (!pc.rev) {
              window16(pc.base, pc.src, limit, w);
            } else {
              u32 t4[4];
              window16(pc.base, pc.src - 15, limit, t4);
              w[0] = __byte_perm(t4[3], 0, 0x0123);
              w[1] = __byte_perm(t4[2], 0, 0x0123);
              w[2] = __byte_perm(t4[1], 0, 0x0123);
              w[3] = __byte_perm(t4[0], 0, 0x0123);
            }
            if (pc.map) {
#pragma unroll
              for (int q = 0; q < 4; q++) {
                const u32 x = w[q];
                w[q] = (u32)lut[x & 0xffu] | ((u32)lut[(x >> 8) & 0xffu] << 8) | ((u32)lut[(x >> 16) & 0xffu] << 16) |
                       ((u32)lut[x >> 24] << 24);
              }
            }
            *reinterpret_cast<uint4 *>(out + o) = make_uint4(w[0], w[1], w[2], w[3]);
            done = true;
          }
        }
      }
      if (!done) s_slow[atomicAdd(&s_nslow, 1u)] = (unsigned short)cidx;
    }
    __syncthreads();
    const u32 nslow = s_nslow;
    for (u32 t = tid; t < nslow; t += 256) {
      const u32 cidx = s_slow[t];
      const u64 o = o0 + (u64)cidx * 16ull;
      const u64 oend = o + 16 < total ? o + 16 : total;
      u32 i;
      int rbeg, rend;
      locate(cidx, i, rbeg, rend);
      u32 w[4] = {0, 0, 0, 0};
      emit_chunk_walk(v, c, rec_of(i), (u32)((int)(cidx * 16u) - rbeg), (long long)rend, (long long)(cidx * 16u),
                      (long long)(oend - o0), lut, in_limit, seq_limit, qual_limit, w, [&](RecOut &ro, long long &re, long long pos) {
                        do {
                          i++;
                          re = s_off[i + 1];
                        } while (pos >= re);
                        ro = rec_of(i);
                      });
      *reinterpret_cast<uint4 *>(out + o) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return;
  }
  for (u32 ch = 0; ch < kEmitChunks; ch++) {
    const u64 o = o0 + ((u64)ch * 256u + tid) * 16ull;
    if (o >= total) return;
    u32 r = search_in(off, s_r[0], s_r[1] + 1, o);
    const u64 oend = o + 16 < total ? o + 16 : total;
    u32 w[4] = {0, 0, 0, 0};
    emit_chunk_walk(v, c, load_rec(v, c, r), (u32)(o - off[r]), (long long)off[r + 1], (long long)o, (long long)oend, lut, in_limit,
                    seq_limit, qual_limit, w, [&](RecOut &ro, long long &re, long long pos) {
                      do {
                        r++;
                        re = (long long)off[r + 1];
                      } while (pos >= re);
                      ro = load_rec(v, c, r);
                    });
    *reinterpret_cast<uint4 *>(out + o) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ---- records whose formatted text IS their input text ("@h\ns\n+\nq\n" / ">h\ns\n" on one line): rmdup, grep and
// seq-with-filters then only compact byte ranges.  k_contig_check counts the kept records that are not of that kind.
__global__ void k_contig_check(RecViews v, EmitCfg c, const u8 *__restrict__ keep, int fastq, u32 in_bytes,
                               unsigned long long *n_bad) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= v.n_rec) return;
  if (keep && !keep[r]) return;
  const u32 no = v.name_off[r], nl = v.name_len[r], so = v.seq_off[r], sl = v.seq_len[r];
  // "<marker>name\n" directly in front of the sequence, and the sequence view is the whole line
  bool ok = no >= 1 && v.in[no - 1] == c.marker && so == no + nl + 1;
  if (fastq) {
    const u32 qo = v.qual_off[r];
    ok = ok && v.qual_len[r] == sl && qo == so + sl + 3 && v.in[so + sl] == '\n' && v.in[so + sl + 1] == '+' &&
         v.in[so + sl + 2] == '\n' && (qo + sl >= in_bytes || v.in[qo + sl] == '\n');
  } else {
    ok = ok && wrap_len(sl, c.width) == sl && (so + sl >= in_bytes || v.in[so + sl] == '\n');
  }
  if (!ok) atomicAdd(n_bad, 1ull);
}

// The CTA's slice of the record table -- output offsets relative to the CTA's first byte, source offsets -- is staged in
// shared memory first (coalesced loads), and a chunk -> record map is built from it (every record with output marks the
// first chunk that starts inside it, a prefix maximum fills the rest), so that a chunk costs three shared-memory loads
// instead of a binary search; a chunk that lies inside one record (94 % of them for 150 bp reads) is one unaligned
// 16-byte window and one store.  Slices of more than kContigCap records (tiny records) read the global arrays.
static const u32 kContigCap = kSliceCap;
static const u32 kContigChunks = 8;  // 16-byte chunks per thread: a CTA writes 32 KiB (the prologue is paid once per CTA)
__global__ void __launch_bounds__(256) k_emit_contig(RecViews v, const u64 *__restrict__ off, u8 *__restrict__ out, u64 total,
                                                     u32 in_bytes, int stage) {
  constexpr u32 NCHUNK = 256 * kContigChunks;  // 16-byte chunks per CTA
  __shared__ u32 s_r[2];
  __shared__ int s_off[kContigCap + 2];   // off[r0 + i] - o0 (negative for a record that starts in front of the CTA)
  __shared__ u32 s_src[kContigCap + 2];   // first input byte of record r0 + i
  __shared__ u32 s_map[NCHUNK];           // chunk -> last record with output that starts at or before the chunk
  __shared__ u32 s_wmax[8];
  __shared__ unsigned short s_slow[NCHUNK];  // chunks that are not one plain window
  __shared__ u32 s_nslow;
  const u32 tid = threadIdx.x;
  if (tid == 0) s_nslow = 0;
  const u64 o0 = (u64)blockIdx.x * NCHUNK * 16ull;
  cta_record_range(off, v.n_rec, o0, total, s_r, (u64)NCHUNK * 16ull);
  const u32 r0 = s_r[0], nr = s_r[1] - s_r[0] + 1;  // records r0 .. r0 + nr - 1; entry nr = end of the last one
  const bool staged = stage && nr <= kContigCap;
  if (staged) {
    for (u32 i = tid; i < nr; i += 256) s_src[i] = v.name_off[r0 + i] - 1u;
    cta_slice_map<kContigChunks>(off, r0, nr, o0, s_off, s_map, s_wmax);
  }
  if (staged) {
    // phase 1: every chunk that lies inside one record issues its five word loads; nothing waits for them yet, so a
    // thread has kEmitChunks x 20 bytes in flight (the kernel is bound by the bytes in flight per SM, not by issue)
#pragma unroll 1
    for (u32 half = 0; half < kContigChunks / 4; half++) {
    u32 ld[4][5], sh[4];
    bool fast[4];
#pragma unroll
    for (u32 ch = 0; ch < 4; ch++) {
      const u32 cidx = (half * 4u + ch) * 256u + tid;
      const u64 o = o0 + (u64)cidx * 16ull;
      fast[ch] = false;
      sh[ch] = 0;
      if (o + 16 <= total) {
        const int ro = (int)(cidx * 16u);
        u32 i = s_map[cidx];
        int rbeg = s_off[i], rend = s_off[i + 1];
        while (ro >= rend) {  // entry 0 of the map may be a record without output
          i++;
          rbeg = rend;
          rend = s_off[i + 1];
        }
        const u64 src = (u64)s_src[i] + (u64)(ro - rbeg);
        const u64 a4 = src & ~3ull;
        if (rend >= ro + 16 && a4 + 20 <= (u64)in_bytes) {  // the whole chunk lies in record i, the window in the input
          fast[ch] = true;
          sh[ch] = (u32)(src & 3ull) * 8u;
          const u32 *wp = reinterpret_cast<const u32 *>(v.in + a4);
#pragma unroll
          for (int q = 0; q < 5; q++) ld[ch][q] = wp[q];
        }
      }
    }
    // phase 2: shift and store.  The other chunks (record boundaries: one or two lanes of every warp, and the ragged
    // end) are put on a list and done afterwards by as many threads side by side -- inside this loop every warp would
    // run the general walk for its one lane.
#pragma unroll
    for (u32 ch = 0; ch < 4; ch++) {
      const u32 cidx = (half * 4u + ch) * 256u + tid;
      const u64 o = o0 + (u64)cidx * 16ull;
      if (fast[ch]) {
        u32 w[4];
#pragma unroll
        for (int q = 0; q < 4; q++) w[q] = __funnelshift_r(ld[ch][q], ld[ch][q + 1], sh[ch]);
        *reinterpret_cast<uint4 *>(out + o) = make_uint4(w[0], w[1], w[2], w[3]);
      } else if (o < total) {
        s_slow[atomicAdd(&s_nslow, 1u)] = (unsigned short)cidx;
      }
    }
    }
    __syncthreads();
    const u32 nslow = s_nslow;
    for (u32 t = tid; t < nslow; t += 256) {
      const u32 cidx = s_slow[t];
      const u64 o = o0 + (u64)cidx * 16ull;
      u32 w[4] = {0, 0, 0, 0};
      const u64 oend = o + 16 < total ? o + 16 : total;
      const int ro = (int)(cidx * 16u), roend = (int)(oend - o0);
      u32 i = s_map[cidx];
      int rbeg = s_off[i], rend = s_off[i + 1];
      int pos = ro;
      while (pos < roend) {
        while (pos >= rend) {  // next record with output (dropped ones have no bytes)
          i++;
          rbeg = rend;
          rend = s_off[i + 1];
        }
        const int seg_end = rend < roend ? rend : roend;
        const u64 src = (u64)s_src[i] + (u64)(pos - rbeg);
        const u32 shift = (u32)(pos - ro), cnt = (u32)(seg_end - pos);
        u32 ww[4];
        window16(v.in, src - shift, (u64)in_bytes, ww);
        merge16(w, ww, shift, cnt);
        pos = seg_end;
      }
      *reinterpret_cast<uint4 *>(out + o) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return;
  }
  for (u32 ch = 0; ch < kContigChunks; ch++) {
    const u64 o = o0 + ((u64)ch * 256u + tid) * 16ull;
    if (o >= total) return;
    u32 w[4] = {0, 0, 0, 0};
    const u64 oend = o + 16 < total ? o + 16 : total;
    u32 r = search_in(off, s_r[0], s_r[1] + 1, o);
    u64 rbeg = off[r], rend = off[r + 1];
    u64 pos = o;
    while (pos < oend) {
      while (pos >= rend) {  // next record with output (dropped ones have no bytes)
        r++;
        rbeg = rend;
        rend = off[r + 1];
      }
      // bytes [pos, seg_end) of the chunk come from record r, whose text starts one byte before its name
      const u64 seg_end = rend < oend ? rend : oend;
      const u64 src = (u64)(v.name_off[r] - 1u) + (pos - rbeg);
      const u32 shift = (u32)(pos - o);                 // first chunk byte this record supplies
      const u32 cnt = (u32)(seg_end - pos);
      // window aligned to the chunk start; src >= pos >= shift, and bytes in front of the record are masked out
      u32 ww[4];
      window16(v.in, src - shift, (u64)in_bytes, ww);
      merge16(w, ww, shift, cnt);
      pos = seg_end;
    }
    *reinterpret_cast<uint4 *>(out + o) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

void out_len(RecViews v, EmitCfg c, const u8 *keep, u32 *out_len_, cudaStream_t s) {
  BSK_LAUNCH_FLAT(k_out_len, (v.n_rec + 1 + 255) / 256, 256, 0, s, v, c, keep, out_len_);
}
void emit(RecViews v, EmitCfg c, const u64 *out_off, u8 *out, u64 total, const u8 *lut, u64 in_limit, u64 seq_limit,
          u64 qual_limit, cudaStream_t s) {
  if (!total) return;
  const u64 per_cta = 256ull * 16 * kEmitChunks;
  // 6 CTAs / SM at 40 registers (124 B of spills) beat 4 at 62: subseq 2.48 -> 2.23 ms per GiB, seq --min-len 3.02 -> 2.64
  // (5: 2.29, 8: 2.25; profiles/r2_experiments.txt).  BSK_EMIT_MB=4 keeps the unconstrained build for A/B runs.
  static const int mb = [] { const char *e = getenv("BSK_EMIT_MB"); return e ? atoi(e) : 6; }();
  const u32 grid = (u32)((total + per_cta - 1) / per_cta);
  if (mb != 4) BSK_LAUNCH(k_emit<6>, grid, 256, 0, s, v, c, out_off, out, total, lut, in_limit, seq_limit, qual_limit);
  else BSK_LAUNCH(k_emit<4>, grid, 256, 0, s, v, c, out_off, out, total, lut, in_limit, seq_limit, qual_limit);
}
void contig_check(RecViews v, EmitCfg c, const u8 *keep, int fastq, u32 in_bytes, u64 *n_bad, cudaStream_t s) {
  if (v.n_rec)
    BSK_LAUNCH_FLAT(k_contig_check, (v.n_rec + 255) / 256, 256, 0, s, v, c, keep, fastq, in_bytes, (unsigned long long *)n_bad);
}
void emit_contig(RecViews v, const u64 *out_off, u8 *out, u64 total, u32 in_bytes, cudaStream_t s) {
  if (!total) return;
  const u64 per_cta = 256ull * 16 * kContigChunks;
  static const int stage = [] { const char *e = getenv("BSK_CONTIG_STAGE"); return e ? atoi(e) : 1; }();  // A/B switch
  BSK_LAUNCH(k_emit_contig, (u32)((total + per_cta - 1) / per_cta), 256, 0, s, v, out_off, out, total, in_bytes, stage);
}

}  // namespace k
}  // namespace bsk
