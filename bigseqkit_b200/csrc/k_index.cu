// k_index.cu -- delimiter scan -> line / record index, record parse, line squeeze.
//
// Replaces, for a whole partition at once, what the reference does one record at a
// time on the CPU:
//   worker.PlainFile(path, ">" | "\n@!\n+") + ReadFixer   bigseqkit/helper.go:148-178,
//                                                          bigseqkit-lib/helper.go:41-66
//   SeqParser.Read                                         bigseqkit-lib/helper.go:219-325
//   parseHeadIDAndDesc                                     bigseqkit-lib/helper.go:329-369
//   seq.GuessAlphabetLessConservatively (call site)        bigseqkit-lib/helper.go:286-291
#include "kernels.h"

namespace bsk {
namespace k {

// ---------------------------------------------------------------- byte masks
__device__ __forceinline__ u32 eq_mask4(u32 w, u32 pat) {
  u32 m = __vcmpeq4(w, pat) & 0x08040201u;
  return (m | (m >> 8) | (m >> 16) | (m >> 24)) & 0xfu;
}

struct PieceMasks {
  u32 nl;  // bit b: byte pos+b is '\n'
  u32 rs;  // bit b: the line starting at pos+b+1 opens a record
};

// Record-start rule (pinned in SURVEY C.1): FASTA '\n' followed by '>'; FASTQ '\n'
// followed by '@' unless the two bytes before that '\n' are "\n+".
__device__ __forceinline__ PieceMasks piece_masks(const u8 *__restrict__ d, u32 n, u32 pos, bool fq, u8 marker) {
  PieceMasks pm;
  pm.nl = 0;
  pm.rs = 0;
  if (pos >= n) return pm;
  u32 w0, w1, w2, w3;
  if (pos + 16 <= n) {
    uint4 v = *reinterpret_cast<const uint4 *>(d + pos);
    w0 = v.x; w1 = v.y; w2 = v.z; w3 = v.w;
  } else {
    u32 w[4] = {0, 0, 0, 0};
    for (u32 b = 0; pos + b < n; b++) w[b >> 2] |= (u32)d[pos + b] << (8 * (b & 3));
    w0 = w[0]; w1 = w[1]; w2 = w[2]; w3 = w[3];
  }
  u32 nl = eq_mask4(w0, 0x0a0a0a0au) | (eq_mask4(w1, 0x0a0a0a0au) << 4) | (eq_mask4(w2, 0x0a0a0a0au) << 8) |
           (eq_mask4(w3, 0x0a0a0a0au) << 12);
  if (pos + 16 > n) nl &= (1u << (n - pos)) - 1u;
  pm.nl = nl;
  u32 m = nl;
  while (m) {
    int b = __ffs((int)m) - 1;
    m &= m - 1;
    u32 i = pos + (u32)b;
    if (i + 1 < n && d[i + 1] == marker) {
      bool hit = true;
      if (fq && i >= 2 && d[i - 2] == '\n' && d[i - 1] == '+') hit = false;
      if (hit) pm.rs |= 1u << b;
    }
  }
  return pm;
}

// ---------------------------------------------------------------- pass 1: counts per 16 KiB tile
__global__ void __launch_bounds__(256) k_index_count(const u8 *__restrict__ d, u32 n, u64 *__restrict__ tile_cnt) {
  __shared__ u32 s_cnt[2];
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const bool fq = d[0] == '@';
  const u8 marker = fq ? '@' : '>';
  const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 base = blockIdx.x * kIndexTile + warp * 2048u;
  u32 cn = 0, cr = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    PieceMasks pm = piece_masks(d, n, base + (u32)j * 512u + lane * 16u, fq, marker);
    cn += (u32)__popc(pm.nl);
    cr += (u32)__popc(pm.rs);
  }
  u32 packed = cn | (cr << 16);
  for (int off = 16; off > 0; off >>= 1) packed += __shfl_xor_sync(0xffffffffu, packed, off);
  if (lane == 0) {
    atomicAdd(&s_cnt[0], packed & 0xffffu);
    atomicAdd(&s_cnt[1], packed >> 16);
  }
  __syncthreads();
  if (threadIdx.x == 0) tile_cnt[blockIdx.x] = (u64)s_cnt[0] | ((u64)s_cnt[1] << 32);
}

// ---------------------------------------------------------------- pass 2: write ls[] and rl[]
__global__ void __launch_bounds__(256) k_index_fill(const u8 *__restrict__ d, u32 n, const u64 *__restrict__ tile_base,
                                                    u32 *__restrict__ ls, u32 *__restrict__ rl) {
  __shared__ u32 s_wtot[8];
  const bool fq = d[0] == '@';
  const u8 marker = fq ? '@' : '>';
  const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 base = blockIdx.x * kIndexTile + warp * 2048u;
  PieceMasks pm[4];
  u32 cnt[4];
  u32 tot = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    pm[j] = piece_masks(d, n, base + (u32)j * 512u + lane * 16u, fq, marker);
    cnt[j] = (u32)__popc(pm[j].nl) | ((u32)__popc(pm[j].rs) << 16);
    tot += cnt[j];
  }
  u32 wt = tot;
  for (int off = 16; off > 0; off >>= 1) wt += __shfl_xor_sync(0xffffffffu, wt, off);
  if (lane == 0) s_wtot[warp] = wt;
  __syncthreads();
  u32 wbase = 0;
  for (u32 w = 0; w < warp; w++) wbase += s_wtot[w];
  const u64 tb = tile_base[blockIdx.x];
  u32 nl_base = (u32)tb + (wbase & 0xffffu);
  u32 rs_base = (u32)(tb >> 32) + (wbase >> 16);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    u32 x = cnt[j];
    for (int off = 1; off < 32; off <<= 1) {
      u32 y = __shfl_up_sync(0xffffffffu, x, off);
      if ((int)lane >= off) x += y;
    }
    const u32 total = __shfl_sync(0xffffffffu, x, 31);
    const u32 excl = x - cnt[j];
    u32 g = nl_base + (excl & 0xffffu);
    u32 q = rs_base + (excl >> 16);
    const u32 pos = base + (u32)j * 512u + lane * 16u;
    u32 m = pm[j].nl;
    while (m) {
      int b = __ffs((int)m) - 1;
      m &= m - 1;
      ls[g + 1] = pos + (u32)b + 1u;
      if ((pm[j].rs >> b) & 1u) {
        rl[q + 1] = g + 1;
        q++;
      }
      g++;
    }
    nl_base += total & 0xffffu;
    rs_base += total >> 16;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ls[0] = 0;
    rl[0] = 0;
  }
}

__global__ void k_index_finish(u32 *ls, u32 *rl, u32 n, u32 n_nl, u32 n_rec, u32 n_lines) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    ls[n_nl + 1] = n + 1;
    rl[n_rec] = n_lines;
  }
}

// ---------------------------------------------------------------- SeqParser.Read, one thread per record
__global__ void k_parse_records(RecIndex ix, RecArrays ra, DevStatus *st) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ix.n_rec) return;
  const u8 *__restrict__ d = ix.in;
  const u32 *__restrict__ ls = ix.ls;
  const u8 marker = ix.fastq ? '@' : '>';
  const u32 l0 = ix.rl[r], l1 = ix.rl[r + 1];
  const u32 hs = ls[l0];
  const u32 mk = (hs < ix.n && d[hs] == marker) ? 1u : 0u;  // ReadFixer prepends the marker when absent
  const u32 h_end = ls[l0 + 1] - 1;
  ra.head_off[r] = hs + mk;
  ra.head_len[r] = h_end - (hs + mk);
  u32 sa = 0, sb = 0, qa = 0, qb = 0;  // line runs [sa,sb) sequence, [qa,qb) quality
  if (l1 - l0 >= 2) {
    if (!ix.fastq) {  // helper.go:240-250: every following line, the unterminated one included
      sa = l0 + 1;
      sb = l1;
    } else {  // helper.go:251-273
      u32 plus = l1;
      for (u32 kk = l0 + 1; kk + 1 < l1; kk++) {
        const u32 a = ls[kk], e = ls[kk + 1] - 1;
        if (e > a && d[a] == '+') { plus = kk; break; }
      }
      sa = l0 + 1;
      if (plus < l1) {
        sb = plus;
        qa = plus + 1;
        qb = l1;
      } else {
        sb = l1 - 1;  // still in sequence mode: the unterminated last segment is dropped
      }
    }
  }
  const u32 slen = sb > sa ? ls[sb] - ls[sa] - (sb - sa) : 0;
  const u32 qlen = qb > qa ? ls[qb] - ls[qa] - (qb - qa) : 0;
  u32 snl = sb - sa, qnl = qb - qa;
  // a run whose bytes all sit on its first line is contiguous in the input
  if (snl > 1 && slen == ls[sa + 1] - 1 - ls[sa]) snl = 1;
  if (qnl > 1 && qlen == ls[qa + 1] - 1 - ls[qa]) qnl = 1;
  ra.seq_line0[r] = sa;
  ra.seq_line1[r] = sb;
  ra.seq_off[r] = snl ? ls[sa] : 0;
  ra.seq_len[r] = slen;
  ra.qual_line0[r] = qa;
  ra.qual_line1[r] = qb;
  ra.qual_off[r] = qnl ? ls[qa] : 0;
  ra.qual_len[r] = ix.fastq ? qlen : 0;
  if ((snl > 1 || qnl > 1) && st->multiline == 0) atomicOr(&st->multiline, 1u);
  if (ix.fastq && slen != qlen) atomicMin((unsigned long long *)&st->err, ((unsigned long long)r << 4) | EK_UNMATCHED);
}

// ---------------------------------------------------------------- multi-line records -> contiguous arenas
// one warp per line; the record of a line is found by binary search in rl[]
__global__ void k_squeeze_lines(RecIndex ix, RecArrays ra, const u32 *__restrict__ seq_aoff,
                                const u32 *__restrict__ qual_aoff, u8 *__restrict__ seq_arena,
                                u8 *__restrict__ qual_arena) {
  const u32 lane = threadIdx.x & 31;
  const u32 line = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (line >= ix.n_lines) return;
  u32 lo = 0, hi = ix.n_rec;  // rl[lo] <= line < rl[hi]
  while (hi - lo > 1) {
    const u32 mid = lo + ((hi - lo) >> 1);
    if (ix.rl[mid] <= line) lo = mid;
    else hi = mid;
  }
  const u32 r = lo;
  const u32 a = ix.ls[line], len = ix.ls[line + 1] - 1 - a;
  if (len == 0) return;
  u8 *dst = nullptr;
  const u32 sa = ra.seq_line0[r], sb = ra.seq_line1[r], qa = ra.qual_line0[r], qb = ra.qual_line1[r];
  if (line >= sa && line < sb) dst = seq_arena + seq_aoff[r] + (a - ix.ls[sa] - (line - sa));
  else if (line >= qa && line < qb && qual_arena) dst = qual_arena + qual_aoff[r] + (a - ix.ls[qa] - (line - qa));
  if (!dst) return;
  const u8 *src = ix.in + a;
  for (u32 i = lane; i < len; i += 32) dst[i] = src[i];
}

// ---------------------------------------------------------------- uniformly wrapped FASTA: arithmetic squeeze
// A record whose sequence lines all have the same width W except the last (<= W) -- what every FASTA writer
// produces -- needs no per-line work: byte q of its squeezed sequence is input byte seq_start + q + q / W.
// k_lines_uniform counts the lines that break that rule (FASTA only: a header is line 0 or a line starting '>').
__global__ void k_lines_uniform(RecIndex ix, unsigned long long *n_bad) {
  const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0 || j >= ix.n_lines) return;
  const u8 *__restrict__ d = ix.in;
  const u32 *__restrict__ ls = ix.ls;
  const u32 a = ls[j], pa = ls[j - 1];
  const bool hdr = a < ix.n && d[a] == '>';
  const bool prev_hdr = j - 1 == 0 || d[pa] == '>';
  if (hdr || prev_hdr) return;  // headers and first sequence lines set no constraint
  const u32 len = ls[j + 1] - 1 - a, plen = a - 1 - pa;
  const bool last = j + 1 >= ix.n_lines || (ls[j + 1] < ix.n && d[ls[j + 1]] == '>');
  if (last ? len > plen : len != plen) atomicAdd(n_bad, 1ull);
}

__device__ __forceinline__ u32 warp_search_u32(const u32 *__restrict__ off, u32 n, u32 o) {  // last r with off[r] <= o
  const u32 lane = threadIdx.x & 31;
  u32 lo = 0, hi = n;
  while (hi - lo > 1) {
    const u32 step = (hi - lo + 32) / 33;
    const u32 idx = lo + (lane + 1) * step;
    const bool le = idx < hi && off[idx] <= o;
    const u32 cnt = (u32)__popc(__ballot_sync(0xffffffffu, le));
    const u32 nhi = lo + (cnt + 1) * step;
    lo += cnt * step;
    if (nhi < hi) hi = nhi;
  }
  return lo;
}

// 16 arena bytes per thread and step (aligned 16-byte stores); each run of bytes inside one input line is one
// unaligned 16-byte window (five aligned word loads + funnel shifts), masked into place
__global__ void __launch_bounds__(256) k_squeeze_uniform(RecIndex ix, RecArrays ra, const u32 *__restrict__ seq_aoff,
                                                         u8 *__restrict__ seq_arena, u32 total) {
  __shared__ u32 s_r[2];
  const u32 cta0 = blockIdx.x * 256u * 16u * 4u;
  if (threadIdx.x < 64) {
    const u32 w = threadIdx.x >> 5;
    u32 o = w == 0 ? cta0 : cta0 + 256u * 16u * 4u - 1u;
    if (o >= total) o = total - 1;
    const u32 r = warp_search_u32(seq_aoff, ix.n_rec, o);
    if ((threadIdx.x & 31) == 0) s_r[w] = r;
  }
  __syncthreads();
  const u8 *__restrict__ in = ix.in;
  for (u32 ch = 0; ch < 4; ch++) {
    const u32 o = cta0 + (ch * 256u + threadIdx.x) * 16u;
    if (o >= total) return;
    u32 lo = s_r[0], hi = s_r[1] + 1;  // seq_aoff[lo] <= o < seq_aoff[hi]
    while (hi - lo > 1) {
      const u32 mid = lo + ((hi - lo) >> 1);
      if (seq_aoff[mid] <= o) lo = mid;
      else hi = mid;
    }
    u32 r = lo;
    u32 rbeg = seq_aoff[r], rend = seq_aoff[r + 1];
    u32 w[4] = {0, 0, 0, 0};
    u32 pos = o;
    const u32 oend = o + 16 < total ? o + 16 : total;
    u32 sstart = 0, W = 1;
    bool have = false;
    while (pos < oend) {
      if (pos >= rend || !have) {
        while (pos >= rend) {  // next record with sequence bytes
          r++;
          rbeg = rend;
          rend = seq_aoff[r + 1];
        }
        const u32 sa = ra.seq_line0[r];
        sstart = ix.ls[sa];
        W = ix.ls[sa + 1] - 1 - sstart;  // width of the record's lines (> 0: the record has sequence bytes)
        have = true;
      }
      const u32 qr = pos - rbeg;
      const u32 line = qr / W, col = qr - line * W;
      u32 cnt = W - col;                                // bytes left on this input line
      if (cnt > rend - pos) cnt = rend - pos;
      if (cnt > oend - pos) cnt = oend - pos;
      const u32 src = sstart + qr + line;
      const u32 shift = pos - o;
      // window aligned to the chunk start (src >= shift: the header line precedes the sequence)
      const u32 s0 = src - shift, a0 = s0 & ~3u, sh = (s0 & 3u) * 8u;
      const u32 *wp = reinterpret_cast<const u32 *>(in + a0);
      u32 x0, x1, x2, x3, x4;
      if (a0 + 20 <= ix.n) { x0 = wp[0]; x1 = wp[1]; x2 = wp[2]; x3 = wp[3]; x4 = wp[4]; }
      else {
        x0 = a0 < ix.n ? wp[0] : 0u; x1 = a0 + 4 < ix.n ? wp[1] : 0u; x2 = a0 + 8 < ix.n ? wp[2] : 0u;
        x3 = a0 + 12 < ix.n ? wp[3] : 0u; x4 = a0 + 16 < ix.n ? wp[4] : 0u;
      }
      const u32 ww[4] = {__funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh), __funnelshift_r(x2, x3, sh),
                         __funnelshift_r(x3, x4, sh)};
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int l = (int)shift - 4 * q, h = (int)(shift + cnt) - 4 * q;  // bytes [l, h) of word q are taken
        if (h <= 0 || l >= 4) continue;
        u32 m = 0xffffffffu;
        if (l > 0) m &= 0xffffffffu << (8 * l);
        if (h < 4) m &= 0xffffffffu >> (8 * (4 - h));
        w[q] |= ww[q] & m;
      }
      pos += cnt;
    }
    if (o + 16 <= total) *reinterpret_cast<uint4 *>(seq_arena + o) = make_uint4(w[0], w[1], w[2], w[3]);
    else for (u32 t = 0; o + t < total; t++) seq_arena[o + t] = (u8)(w[t >> 2] >> (8 * (t & 3)));
  }
}

// ---------------------------------------------------------------- alphabet guess on record 0
__global__ void k_guess_alphabet(RecViews v, const u8 *__restrict__ class_mask, u32 limit, DevStatus *st) {
  if (v.n_rec == 0) return;
  u32 len = v.seq_len[0];
  if (limit > 0 && len > limit) len = limit;
  const u8 *s = v.seqb + v.seq_off[0];
  u32 m = 0xffffffffu;
  for (u32 i = threadIdx.x; i < len; i += blockDim.x) m &= class_mask[s[i]];
  if (m != 0xffffffffu) atomicAnd(&st->guess_mask, m);
  if (threadIdx.x == 0) st->guess_len = len;
}

// ---------------------------------------------------------------- parseHeadIDAndDesc
__global__ void k_id_desc(RecViews v, int id_ncbi, u32 *id_off, u32 *id_len, u32 *desc_off, u32 *desc_len) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= v.n_rec) return;
  const u32 ho = v.name_off[r], e = v.name_len[r];
  const u8 *h = v.in + ho;
  u32 io = ho, il = e, dof = ho + e, dl = 0;
  if (id_ncbi) {  // leftmost match of \|([^\|]+)\|<space>  (bigseqkit/helper.go:97-100)
    for (u32 i = 0; i < e; i++) {
      if (h[i] != '|') continue;
      u32 j = i + 1;
      while (j < e && h[j] != '|') j++;
      if (j < e && j > i + 1 && j + 1 < e && h[j + 1] == ' ') {
        io = ho + i + 1;
        il = j - i - 1;
        break;
      }
    }
  } else {
    u32 i = e;
    for (u32 t = 0; t < e; t++)
      if (h[t] == ' ') { i = t; break; }
    if (i == e || i == 0) {
      i = e;
      for (u32 t = 0; t < e; t++)
        if (h[t] == '\t') { i = t; break; }
      if (i == 0) i = e;
    }
    if (i < e) {
      u32 j = i + 1;
      while (j < e && (h[j] == ' ' || h[j] == '\t')) j += 2;  // sic: helper.go:334-339 advances twice
      il = i;
      if (j < e) {
        dof = ho + j;
        dl = e - j;
      }
    }
  }
  id_off[r] = io;
  id_len[r] = il;
  if (desc_off) {
    desc_off[r] = dof;
    desc_len[r] = dl;
  }
}

// ---------------------------------------------------------------- Alphabet.IsValid on the first `limit` bytes
__global__ void k_validate_seq(RecViews v, const u8 *__restrict__ valid, u32 limit, DevStatus *st) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= v.n_rec) return;
  u32 len = v.seq_len[r];
  if (limit > 0 && len > limit) len = limit;
  const u8 *s = v.seqb + v.seq_off[r];
  for (u32 i = 0; i < len; i++)
    if (!valid[s[i]]) {
      atomicMin((unsigned long long *)&st->err, ((unsigned long long)r << 4) | EK_VALIDATE);
      return;
    }
}

// ---------------------------------------------------------------- launchers
void index_count(const u8 *in, u32 n, u64 *tile_cnt, u32 n_tiles, cudaStream_t s) {
  if (n_tiles) BSK_LAUNCH(k_index_count, n_tiles, 256, 0, s, in, n, tile_cnt);
}
void index_fill(const u8 *in, u32 n, const u64 *tile_base, u32 *ls, u32 *rl, u32 n_tiles, cudaStream_t s) {
  if (n_tiles) BSK_LAUNCH(k_index_fill, n_tiles, 256, 0, s, in, n, tile_base, ls, rl);
}
void index_finish(u32 *ls, u32 *rl, u32 n, u32 n_nl, u32 n_rec, u32 n_lines, cudaStream_t s) {
  BSK_LAUNCH_FLAT(k_index_finish, 1, 1, 0, s, ls, rl, n, n_nl, n_rec, n_lines);
}
void parse_records(RecIndex ix, RecArrays ra, DevStatus *st, cudaStream_t s) {
  if (ix.n_rec) BSK_LAUNCH_FLAT(k_parse_records, (ix.n_rec + 255) / 256, 256, 0, s, ix, ra, st);
}
void squeeze_lines(RecIndex ix, RecArrays ra, const u32 *seq_aoff, const u32 *qual_aoff, u8 *seq_arena, u8 *qual_arena,
                   cudaStream_t s) {
  if (!ix.n_lines) return;
  const u64 threads = (u64)ix.n_lines * 32;
  BSK_LAUNCH_FLAT(k_squeeze_lines, (u32)((threads + 255) / 256), 256, 0, s, ix, ra, seq_aoff, qual_aoff, seq_arena,
                  qual_arena);
}
void lines_uniform(RecIndex ix, u64 *n_bad, cudaStream_t s) {
  if (ix.n_lines > 1) BSK_LAUNCH_FLAT(k_lines_uniform, (ix.n_lines + 255) / 256, 256, 0, s, ix, (unsigned long long *)n_bad);
}
void squeeze_uniform(RecIndex ix, RecArrays ra, const u32 *seq_aoff, u8 *seq_arena, u32 total, cudaStream_t s) {
  if (!total) return;
  const u32 per_cta = 256u * 16u * 4u;
  BSK_LAUNCH(k_squeeze_uniform, (total + per_cta - 1) / per_cta, 256, 0, s, ix, ra, seq_aoff, seq_arena, total);
}
void guess_alphabet(RecViews v, const u8 *class_mask, u32 limit, DevStatus *st, cudaStream_t s) {
  BSK_LAUNCH_FLAT(k_guess_alphabet, 1, 256, 0, s, v, class_mask, limit, st);
}
void id_desc(RecViews v, int id_ncbi, u32 *id_off, u32 *id_len, u32 *desc_off, u32 *desc_len, cudaStream_t s) {
  if (v.n_rec) BSK_LAUNCH_FLAT(k_id_desc, (v.n_rec + 255) / 256, 256, 0, s, v, id_ncbi, id_off, id_len, desc_off, desc_len);
}
void validate_seq(RecViews v, const u8 *valid, u32 limit, DevStatus *st, cudaStream_t s) {
  if (v.n_rec) BSK_LAUNCH_FLAT(k_validate_seq, (v.n_rec + 255) / 256, 256, 0, s, v, valid, limit, st);
}

}  // namespace k
}  // namespace bsk
