// k_fused.cu -- single-pass tile kernels for SHORT records (reads, CDS): the whole hot path of
// `seq` in one launch.
//
//   PlainFile split + ReadFixer     bigseqkit/helper.go:148-178, bigseqkit-lib/helper.go:41-66
//   SeqParser.Read                  bigseqkit-lib/helper.go:219-325
//   SeqTransform.Call               bigseqkit-lib/seq.go:81-269
//   FileStore framing               bigseqkit-lib/helper.go:441-451
//
// One CTA owns the records that START inside its 16 KiB input tile.  The tile (plus a halo that
// holds the tail of the last owned record) is staged in shared memory with 16-byte loads, newlines
// and record starts are found there, every owned record is parsed by one thread, output sizes are
// scanned inside the CTA and chained across CTAs with a decoupled look-back, and the records are
// assembled in a shared-memory output tile that is flushed with aligned 16-byte stores.  HBM sees
// the input once and the output once (2N algorithmic bytes for seq -r -p on FASTQ).
//
// Anything the tile scheme cannot represent (a record longer than the halo, multi-line FASTQ,
// malformed records, >1024 records or >3072 lines per tile) raises a flag and the caller re-runs
// the block on the general index/parse/emit path, which also produces the reference's error text.
#include "kernels.h"

namespace bsk {
namespace k {

static const u32 FT = 16384;    // tile bytes
static const u32 FH = 4096;     // halo bytes (longest record the fused path accepts, roughly)
static const u32 FPRE = 16;     // look-behind bytes kept in front of the tile
static const u32 FNL = 3072;    // newline slots per region
static const u32 FREC = 1024;   // owned records per tile
static const u32 FOC = 24576;   // output staging bytes per round
static const u32 FTHREADS = 256;

struct FusedSmem {
  u8 in[FPRE + FT + FH + 32];
  u8 out[FOC + 48];
  u8 lut[256];
  u16 nl[FNL + 2];    // newline positions (region coordinates)
  u16 rs[FREC + 2];   // newline index in front of every record start found in the region
  u16 r_sp[FREC];     // record start
  u16 r_ho[FREC], r_hl[FREC], r_so[FREC], r_sl[FREC], r_qo[FREC];
  u32 r_olen[FREC];   // output bytes (0 = dropped)
  u32 r_ooff[FREC + 1];
  u32 wsum[16];
  u32 n_nl, n_rs, n_own, fallback;
  u32 tile;
  u32 keep_total;
  unsigned long long obase;
  u32 rec_base, keep_base;
};

struct TileState {
  unsigned long long agg_bytes, incl_bytes;
  u32 agg_rec, incl_rec, agg_keep, incl_keep;
  u32 status, pad;
};

struct FusedSeqArgs {
  const u8 *in;
  u32 n;
  u8 *out;
  u64 *elem_off;  // may be null
  const u8 *lut;  // 256-byte map of sequence bytes (identity when no transform)
  TileState *tiles;
  u32 *ticket;
  DevStatus *st;  // counters[0] = fallback flag, [1] = total out bytes, [2] = records, [3] = kept
  u8 marker, print_name, print_seq, print_qual, plus_line, reverse, only_id, fastq;
  u32 width;
  int min_len, max_len;
};

__device__ __forceinline__ u32 f_wrap_len(u32 l, u32 w) { return (w < 1 || l == 0) ? l : l + (l - 1) / w; }

__device__ __forceinline__ u32 block_excl_scan(u32 v, u32 *wsum, u32 &total) {
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 x = v;
  for (int off = 1; off < 32; off <<= 1) {
    const u32 y = __shfl_up_sync(0xffffffffu, x, off);
    if ((int)lane >= off) x += y;
  }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  u32 base = 0, tot = 0;
  for (u32 w = 0; w < (blockDim.x >> 5); w++) {
    const u32 s = wsum[w];
    if (w < warp) base += s;
    tot += s;
  }
  __syncthreads();
  total = tot;
  return base + x - v;
}

// stage [t0 - FPRE, t0 + FT + FH) into sm.in; bytes outside the file read as '\n'
__device__ __forceinline__ void fused_load(FusedSmem &sm, const u8 *__restrict__ in, u32 n, u32 t0) {
  const u32 chunks = (FPRE + FT + FH) / 16;
  for (u32 c = threadIdx.x; c < chunks; c += blockDim.x) {
    const long long g = (long long)t0 - FPRE + (long long)c * 16;
    uint4 v;
    if (g >= 0 && g + 16 <= (long long)n) {
      v = *reinterpret_cast<const uint4 *>(in + g);
    } else {
      u32 w[4] = {0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au};
      for (int b = 0; b < 16; b++) {
        const long long p = g + b;
        if (p >= 0 && p < (long long)n) w[b >> 2] = (w[b >> 2] & ~(0xffu << (8 * (b & 3)))) | ((u32)in[p] << (8 * (b & 3)));
      }
      v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4 *>(sm.in + c * 16) = v;
  }
}

// newline + record-start lists of the region [0, lim) (region coordinate p <-> global t0 + p)
__device__ __forceinline__ void fused_scan_lines(FusedSmem &sm, u32 n, u32 t0, bool fq) {
  const u8 marker = fq ? '@' : '>';
  const u32 lim = (n - t0 < FT + FH) ? n - t0 : FT + FH;  // valid bytes in the region
  const bool eof_in_region = (n - t0) <= FT + FH;
  const u8 *d = sm.in + FPRE;
  // each thread owns (FT+FH)/256 = 80 consecutive bytes
  const u32 per = (FT + FH) / FTHREADS;
  const u32 p0 = threadIdx.x * per;
  u32 cnt_nl = 0, cnt_rs = 0;
  for (u32 i = 0; i < per; i++) {
    const u32 p = p0 + i;
    if (p < lim && d[p] == '\n') {
      cnt_nl++;
      const u32 g = t0 + p;  // global position of the newline
      if (g + 1 < n && d[p + 1] == marker && !(fq && g >= 2 && d[p - 2] == '\n' && d[p - 1] == '+')) cnt_rs++;
    }
  }
  u32 tot;
  const u32 packed = cnt_nl | (cnt_rs << 16);
  const u32 ex = block_excl_scan(packed, sm.wsum, tot);
  u32 k_nl = ex & 0xffffu, k_rs = ex >> 16;
  const u32 tot_nl = tot & 0xffffu, tot_rs = tot >> 16;
  const bool virt_nl = eof_in_region && lim > 0 && d[lim - 1] != '\n';  // unterminated last line
  if (tot_nl + (virt_nl ? 1 : 0) > FNL || tot_rs > FREC) {
    if (threadIdx.x == 0) sm.fallback = 1;
  } else {
    for (u32 i = 0; i < per; i++) {
      const u32 p = p0 + i;
      if (p < lim && d[p] == '\n') {
        sm.nl[k_nl] = (u16)p;
        const u32 g = t0 + p;
        if (g + 1 < n && d[p + 1] == marker && !(fq && g >= 2 && d[p - 2] == '\n' && d[p - 1] == '+')) sm.rs[k_rs++] = (u16)k_nl;
        k_nl++;
      }
    }
    if (threadIdx.x == 0) {
      u32 c = tot_nl;
      if (virt_nl) sm.nl[c++] = (u16)lim;
      sm.n_nl = c;
      sm.n_rs = tot_rs;
    }
  }
  __syncthreads();
}

// parseHeadIDAndDesc default regexp on a header held in shared memory: returns the ID length
__device__ __forceinline__ u32 f_id_len(const u8 *h, u32 e) {
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

__global__ void __launch_bounds__(FTHREADS) k_seq_fused(FusedSeqArgs a) {
  BSK_DYN_SMEM(FusedSmem, smp);
  FusedSmem &sm = *smp;
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    sm.tile = atomicAdd(a.ticket, 1u);
    sm.fallback = 0;
    sm.n_nl = sm.n_rs = sm.n_own = 0;
  }
  sm.lut[threadIdx.x] = a.lut[threadIdx.x];
  __syncthreads();
  const u32 tile = sm.tile;
  const u32 t0 = tile * FT;
  const bool fq = a.fastq != 0;
  fused_load(sm, a.in, a.n, t0);
  __syncthreads();
  fused_scan_lines(sm, a.n, t0, fq);

  // ---- owned records: tile 0 owns the record at offset 0; others own starts in [t0, t0 + FT)
  const u8 *d = sm.in + FPRE;
  const u32 lim = (a.n - t0 < FT + FH) ? a.n - t0 : FT + FH;
  const bool eof_in_region = (a.n - t0) <= FT + FH;
  const u32 n_nl = sm.n_nl, n_rs = sm.n_rs;
  // a record that starts on the tile's first byte is announced by a newline in the look-behind, which the scan
  // does not cover: tile 0 always owns offset 0, later tiles own it when the record-start rule holds at t0
  u32 first0 = tile == 0 ? 1u : 0u;
  if (tile > 0 && d[-1] == '\n' && d[0] == (fq ? '@' : '>') && !(fq && d[-3] == '\n' && d[-2] == '+')) first0 = 1u;
  // number of owned starts among rs[]: start position nl[rs[i]] + 1 < FT
  u32 n_own_rs = 0;
  if (!sm.fallback) {
    u32 lo = 0, hi = n_rs;  // first i with start >= FT
    while (lo < hi) {
      const u32 mid = (lo + hi) >> 1;
      if ((u32)sm.nl[sm.rs[mid]] + 1u < FT) lo = mid + 1;
      else hi = mid;
    }
    n_own_rs = lo;
  }
  const u32 n_own = sm.fallback ? 0 : n_own_rs + first0;
  u32 my_bad = 0;
  for (u32 r = threadIdx.x; r < n_own; r += blockDim.x) {
    // record r: start newline index ka (newline before the start, -1 for the file start), end newline index kb
    const int idx = (int)r - (int)first0;  // index into rs[], -1 = file start
    const u32 sp = idx < 0 ? 0u : (u32)sm.nl[sm.rs[idx]] + 1u;
    const u32 a_nl = idx < 0 ? 0u : (u32)sm.rs[idx] + 1u;  // first newline at or after sp
    u32 b_nl;                                              // newline that ends the record
    if ((u32)(idx + 1) < n_rs) b_nl = sm.rs[idx + 1];
    else if (eof_in_region && n_nl > 0) b_nl = n_nl - 1;
    else { my_bad = 1; continue; }  // record runs past the halo
    if (b_nl < a_nl) {
      // the header line itself is the last line of the region (no newline inside the record)
      if (!(eof_in_region)) { my_bad = 1; continue; }
    }
    const u32 mk = (sp < lim && d[sp] == (fq ? '@' : '>')) ? 1u : 0u;
    const u32 h_end = sm.nl[a_nl];
    u32 ho = sp + mk, hl = h_end - ho, so = 0, sl = 0, qo = 0;
    const u32 nlines = b_nl - a_nl + 1;  // lines of the record (header included)
    if (fq) {
      if (nlines != 4) { my_bad = 1; continue; }
      so = h_end + 1;
      sl = sm.nl[a_nl + 1] - so;
      const u32 pl = sm.nl[a_nl + 1] + 1;
      if (!(sm.nl[a_nl + 2] > pl && d[pl] == '+')) { my_bad = 1; continue; }
      if (sl > 0 && d[so] == '+') { my_bad = 1; continue; }  // a sequence line starting with '+' flips the parser
      qo = sm.nl[a_nl + 2] + 1;
      const u32 ql = sm.nl[a_nl + 3] - qo;
      if (ql != sl) { my_bad = 1; continue; }  // the general path reports the error text
    } else {
      if (nlines >= 2) {
        so = h_end + 1;
        // sequence bytes = span minus the newlines between the lines (squeezed later)
        sl = (sm.nl[b_nl] - so) - (nlines - 2);
      }
      qo = nlines;  // FASTA: stash the line count for the squeeze
    }
    // --- output size (seq.go:133-163, 241-259)
    bool keep = true;
    if (a.min_len > 0 && (int)sl < a.min_len) keep = false;
    if (a.max_len > 0 && (int)sl > a.max_len) keep = false;
    u32 nhl = hl;
    if (a.only_id && a.print_name) nhl = f_id_len(d + ho, hl);
    u32 ol = 0;
    if (keep) {
      if (a.print_name) ol += (a.marker ? 1u : 0u) + nhl + 1u;
      if (a.print_seq) ol += f_wrap_len(sl, a.width) + 1u;
      if (a.print_qual) ol += (a.plus_line ? 2u : 0u) + sl + 1u;
    }
    sm.r_sp[r] = (u16)(a_nl);  // first newline index of the record (used by the squeeze)
    sm.r_ho[r] = (u16)ho;
    sm.r_hl[r] = (u16)nhl;
    sm.r_so[r] = (u16)so;
    sm.r_sl[r] = (u16)sl;
    sm.r_qo[r] = (u16)qo;
    sm.r_olen[r] = ol;
  }
  if (my_bad) sm.fallback = 1;
  __syncthreads();
  if (sm.fallback) {
    // publish an empty tile so that successors do not wait for ever, raise the flag, leave
    if (threadIdx.x == 0) {
      atomicAdd((unsigned long long *)&a.st->counters[0], 1ull);
      TileState *ts = a.tiles + tile;
      ts->agg_bytes = 0; ts->agg_rec = 0; ts->agg_keep = 0;
      ts->incl_bytes = 0; ts->incl_rec = 0; ts->incl_keep = 0;
      __threadfence();
      atomicExch(&ts->status, 3u);  // poisoned
    }
    return;
  }

  // ---- FASTA: squeeze multi-line sequences in place (warp per record)
  if (!fq) {
    for (u32 r = warp; r < n_own; r += (blockDim.x >> 5)) {
      const u32 nlines = sm.r_qo[r];
      if (nlines <= 2) continue;
      const u32 a_nl = sm.r_sp[r];
      u32 dst = (u32)sm.nl[a_nl + 1];  // end of the first sequence line
      for (u32 j = 2; j < nlines; j++) {
        const u32 s0 = (u32)sm.nl[a_nl + j - 1] + 1u, len = (u32)sm.nl[a_nl + j] - s0;
        for (u32 i = 0; i < len; i += 32) {
          u8 c = 0;
          if (i + lane < len) c = sm.in[FPRE + s0 + i + lane];
          __syncwarp();
          if (i + lane < len) sm.in[FPRE + dst + i + lane] = c;
          __syncwarp();
        }
        dst += len;
      }
    }
    __syncthreads();
  }

  // ---- output offsets inside the tile, totals, chained scan across tiles
  u32 run = 0;
  for (u32 base = 0; base < n_own; base += blockDim.x) {
    const u32 r = base + threadIdx.x;
    const u32 v = r < n_own ? sm.r_olen[r] : 0;
    u32 tot;
    const u32 ex = block_excl_scan(v, sm.wsum, tot);
    if (r < n_own) sm.r_ooff[r] = run + ex;
    run += tot;
  }
  u32 kept = 0;
  for (u32 r = threadIdx.x; r < n_own; r += blockDim.x) kept += sm.r_olen[r] ? 1u : 0u;
  u32 kt;
  const u32 kex = block_excl_scan(kept, sm.wsum, kt);
  (void)kex;
  if (threadIdx.x == 0) {
    sm.r_ooff[n_own] = run;
    sm.keep_total = kt;
    TileState *ts = a.tiles + tile;
    unsigned long long ex_b = 0;
    u32 ex_r = 0, ex_k = 0;
    if (tile > 0) {
      ts->agg_bytes = run;
      ts->agg_rec = n_own;
      ts->agg_keep = kt;
      __threadfence();
      atomicExch(&ts->status, 1u);
      int j = (int)tile - 1;
      for (;;) {
        TileState *p = a.tiles + j;
        u32 s;
        while ((s = atomicAdd(&p->status, 0u)) == 0) __nanosleep(40);
        __threadfence();
        if (s == 3u) { sm.fallback = 1; break; }
        if (s == 2u) {
          ex_b += *(volatile unsigned long long *)&p->incl_bytes;
          ex_r += *(volatile u32 *)&p->incl_rec;
          ex_k += *(volatile u32 *)&p->incl_keep;
          break;
        }
        ex_b += *(volatile unsigned long long *)&p->agg_bytes;
        ex_r += *(volatile u32 *)&p->agg_rec;
        ex_k += *(volatile u32 *)&p->agg_keep;
        j--;
      }
    }
    ts->incl_bytes = ex_b + run;
    ts->incl_rec = ex_r + n_own;
    ts->incl_keep = ex_k + kt;
    __threadfence();
    atomicExch(&ts->status, sm.fallback ? 3u : 2u);
    sm.obase = ex_b;
    sm.rec_base = ex_r;
    sm.keep_base = ex_k;
    if ((tile + 1) * FT >= a.n || (tile + 1) == (a.n + FT - 1) / FT) {  // last tile: totals for the host
      a.st->counters[1] = ex_b + run;
      a.st->counters[2] = ex_r + n_own;
      a.st->counters[3] = ex_k + kt;
    }
  }
  __syncthreads();
  if (sm.fallback) return;
  const unsigned long long obase = sm.obase;

  // ---- element offsets of the kept records
  if (a.elem_off) {
    u32 krun = 0;
    for (u32 base = 0; base < n_own; base += blockDim.x) {
      const u32 r = base + threadIdx.x;
      const u32 v = (r < n_own && sm.r_olen[r]) ? 1u : 0u;
      u32 tot;
      const u32 ex = block_excl_scan(v, sm.wsum, tot);
      if (v) a.elem_off[sm.keep_base + krun + ex] = obase + sm.r_ooff[r];
      krun += tot;
    }
  }

  // ---- assemble in shared memory, flush with aligned stores; rounds of FOC output bytes
  const u32 total = run;
  const u32 mis = (u32)(obase & 15ull);  // staging index i <-> global byte (obase - mis) + i
  for (u32 rbase = 0; rbase < total || (total == 0 && rbase == 0); rbase += FOC - 16) {
    if (total == 0) break;
    const u32 rend = (total - rbase < FOC - 16) ? total : rbase + (FOC - 16);
    // records overlapping [rbase, rend): warp per record
    for (u32 r = warp; r < n_own; r += (blockDim.x >> 5)) {
      const u32 ol = sm.r_olen[r];
      if (!ol) continue;
      const u32 oo = sm.r_ooff[r];
      if (oo >= rend || oo + ol <= rbase) continue;
      const u32 hl = sm.r_hl[r], sl = sm.r_sl[r];
      const u8 *hsrc = sm.in + FPRE + sm.r_ho[r];
      const u8 *ssrc = sm.in + FPRE + sm.r_so[r];
      const u8 *qsrc = sm.in + FPRE + sm.r_qo[r];
      // p = position inside the record's output; write iff rbase <= oo + p < rend
      u32 p = 0;
#define F_PUT(pos, byte)                                                          \
  do {                                                                            \
    const u32 gp_ = oo + (pos);                                                   \
    if (gp_ >= rbase && gp_ < rend) sm.out[mis + gp_ - rbase] = (byte);           \
  } while (0)
      if (a.print_name) {
        if (a.marker) {
          if (lane == 0) F_PUT(p, a.marker);
          p += 1;
        }
        for (u32 k2 = lane; k2 < hl; k2 += 32) F_PUT(p + k2, hsrc[k2]);
        p += hl;
        if (lane == 0) F_PUT(p, (u8)'\n');
        p += 1;
      }
      if (a.print_seq) {
        const u32 wl = f_wrap_len(sl, a.width);
        if (a.width == 0) {
          for (u32 k2 = lane; k2 < sl; k2 += 32) F_PUT(p + k2, sm.lut[ssrc[a.reverse ? sl - 1 - k2 : k2]]);
        } else {
          const u32 w1 = a.width + 1;
          for (u32 k2 = lane; k2 < wl; k2 += 32) {
            const u32 line = k2 / w1, col = k2 - line * w1;
            u8 b = '\n';
            if (col != a.width) {
              const u32 j = line * a.width + col;
              b = sm.lut[ssrc[a.reverse ? sl - 1 - j : j]];
            }
            F_PUT(p + k2, b);
          }
        }
        p += wl;
        if (lane == 0) F_PUT(p, (u8)'\n');
        p += 1;
      }
      if (a.print_qual) {
        if (a.plus_line) {
          if (lane == 0) { F_PUT(p, (u8)'+'); F_PUT(p + 1, (u8)'\n'); }
          p += 2;
        }
        for (u32 k2 = lane; k2 < sl; k2 += 32) F_PUT(p + k2, qsrc[a.reverse ? sl - 1 - k2 : k2]);
        p += sl;
        if (lane == 0) F_PUT(p, (u8)'\n');
        p += 1;
      }
#undef F_PUT
    }
    __syncthreads();
    // flush staging [mis, mis + (rend - rbase)) -> global [obase + rbase, obase + rend)
    {
      const u32 nbytes = rend - rbase;
      u8 *gdst = a.out + (obase - mis) + rbase;  // 16-byte aligned
      const u32 first_full = mis ? 16u : 0u;     // first fully owned 16-byte chunk
      const u32 end_idx = mis + nbytes;
      const u32 last_full = end_idx & ~15u;
      if (last_full > first_full) {
        for (u32 i = first_full + threadIdx.x * 16; i < last_full; i += blockDim.x * 16)
          *reinterpret_cast<uint4 *>(gdst + i) = *reinterpret_cast<const uint4 *>(sm.out + i);
      }
      // ragged head and tail bytes
      if (mis) {
        const u32 he = end_idx < 16u ? end_idx : 16u;
        for (u32 i = mis + threadIdx.x; i < he; i += blockDim.x) gdst[i] = sm.out[i];
      }
      const u32 ts_ = last_full > first_full ? last_full : (mis ? (end_idx < 16u ? end_idx : 16u) : 0u);
      for (u32 i = ts_ + threadIdx.x; i < end_idx; i += blockDim.x) gdst[i] = sm.out[i];
    }
    __syncthreads();
    if (rend == total) break;
  }
}

size_t fused_smem_bytes() { return sizeof(FusedSmem) + 16; }
size_t fused_tile_state_bytes(u32 n) { return (size_t)((n + FT - 1) / FT + 1) * sizeof(TileState); }
u32 fused_max_records(u32 n) { return ((n + FT - 1) / FT + 1) * FREC; }

void seq_fused(const u8 *in, u32 n, u8 *out, u64 *elem_off, const u8 *lut, void *tile_state, u32 *ticket, DevStatus *st,
               EmitCfg cfg, int only_id, int fastq, int min_len, int max_len, cudaStream_t s) {
  FusedSeqArgs a;
  a.in = in;
  a.n = n;
  a.out = out;
  a.elem_off = elem_off;
  a.lut = lut;
  a.tiles = static_cast<TileState *>(tile_state);
  a.ticket = ticket;
  a.st = st;
  a.marker = cfg.marker;
  a.print_name = cfg.print_name;
  a.print_seq = cfg.print_seq;
  a.print_qual = cfg.print_qual;
  a.plus_line = cfg.plus_line;
  a.reverse = cfg.reverse;
  a.only_id = (u8)only_id;
  a.fastq = (u8)fastq;
  a.width = cfg.width;
  a.min_len = min_len;
  a.max_len = max_len;
  const u32 n_tiles = (n + FT - 1) / FT;
#ifndef BSK_EMU
  // the opt-in to > 48 KiB of dynamic shared memory is per device (a process may hold ctxs on several GPUs)
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaFuncSetAttribute(k_seq_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem_bytes());
    attr_set[dev] = true;
  }
#endif
  BSK_LAUNCH(k_seq_fused, n_tiles, FTHREADS, fused_smem_bytes(), s, a);
}

}  // namespace k
}  // namespace bsk
