// ops_translate.cu -- translate: 1-6 frame translation with an NCBI genetic code.
//
//   Translate.Before / Call        bigseqkit-lib/translate.go:33-145
//   Seq.Translate, seq.CodonTables (bio v0.7.0; ambiguous-codon rule pinned by bigseqkit-cli/translate.go:42-52)
// Pinned semantics: one element per (record, frame): ">header\n" + wrapped protein.
//
// Device layout: protein arena = all (record, frame) proteins back to back; the record
// formatter of k_emit.cu then treats every (record, frame) pair as a FASTA record whose
// "sequence" lives in that arena.
#include <algorithm>
#include <cstring>

#include "engine.h"
#include "gcode_tables.h"
#include "prims.h"

namespace bsk {

struct TrCfg {
  u32 nf;
  int frames[6 * 11];  // up to 64 frames in the reference; we cap at 66
  int allow_unknown, init_m, clean, trim;
};

// untrimmed protein length per (record, frame); records shorter than 3 nt are an error
__global__ void k_tr_len(RecViews v, TrCfg c, u32 *__restrict__ plen, u32 *__restrict__ ppad, DevStatus *st) {
  const u32 e = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 n_el = v.n_rec * c.nf;
  if (e > n_el) return;
  if (e == n_el) { plen[e] = 0; ppad[e] = 0; return; }
  const u32 r = e / c.nf, fi = e - r * c.nf;
  const u32 l = v.seq_len[r];
  if (l < 3) {
    plen[e] = 0;
    ppad[e] = 0;
    atomicMin((unsigned long long *)&st->err, ((unsigned long long)r << 4) | EK_TOO_SHORT);
    return;
  }
  const int f = c.frames[fi];
  const u32 start = (u32)((f < 0 ? -f : f) - 1);
  plen[e] = (l - start) / 3;
  ppad[e] = ((l - start) / 3 + 15u) & ~15u;  // arena slots are padded to 16 amino acids: one element per 16-byte chunk
}

// One thread per 16 consecutive amino acids of the arena (one aligned 16-byte store); the CTA finds the range of
// (record, frame) elements it covers once with two warp-wide searches, the code tables live in shared memory.
static const u32 kTrAA = 16;

__device__ __forceinline__ u32 tr_warp_search(const u64 *__restrict__ off, u32 n, u64 o) {  // last e with off[e] <= o
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

__global__ void __launch_bounds__(256) k_translate(RecViews v, TrCfg c, const u64 *__restrict__ poff, u64 total,
                                                   const u8 *__restrict__ code_fwd, const u8 *__restrict__ code_rev,
                                                   const u8 *__restrict__ lut, u8 *__restrict__ prot, DevStatus *st,
                                                   u64 seq_limit, const u32 *__restrict__ plen) {
  __shared__ u8 s_fwd[256], s_rev[256];
  __align__(16) __shared__ u8 s_lut[4096];
  __shared__ u32 s_e[2];
  // shared code tables: IUPAC mask 1..15 in the low nibble; gap -> 0x10, not a nucleotide -> 0x20 (low nibble 0), so
  // that one OR over a chunk's codes tells whether any codon needs the careful path
  {
    const u8 a = code_fwd[threadIdx.x], b = code_rev[threadIdx.x];
    s_fwd[threadIdx.x] = a == 16 ? 0x10 : (a == 0 ? 0x20 : a);
    s_rev[threadIdx.x] = b == 16 ? 0x10 : (b == 0 ? 0x20 : b);
  }
  if (((size_t)lut & 15) == 0) {  // 4096 codon entries: one 16-byte copy per thread
    reinterpret_cast<uint4 *>(s_lut)[threadIdx.x] = reinterpret_cast<const uint4 *>(lut)[threadIdx.x];
  } else {
    for (u32 i = threadIdx.x; i < 4096; i += 256) s_lut[i] = lut[i];
  }
  const u32 n_el = v.n_rec * c.nf;
  const u64 cta0 = (u64)blockIdx.x * 256ull * kTrAA;
  if (threadIdx.x < 64) {
    const u32 w = threadIdx.x >> 5;
    u64 o = w == 0 ? cta0 : cta0 + 256ull * kTrAA - 1;
    if (o >= total) o = total - 1;
    const u32 e = tr_warp_search(poff, n_el, o);
    if ((threadIdx.x & 31) == 0) s_e[w] = e;
  }
  __syncthreads();
  const u64 a0 = cta0 + (u64)threadIdx.x * kTrAA;
  if (a0 >= total) return;
  u32 lo = s_e[0], hi = s_e[1] + 1;  // poff[lo] <= a0 < poff[hi]
  while (hi - lo > 1) {
    const u32 mid = lo + ((hi - lo) >> 1);
    if (poff[mid] <= a0) lo = mid;
    else hi = mid;
  }
  u32 e = lo;
  const u32 j = (u32)(a0 - poff[e]);  // slots are padded to 16: the whole chunk belongs to element e
  const u32 r = e / c.nf;
  const int f = c.frames[e - r * c.nf];
  const u32 start = (u32)((f < 0 ? -f : f) - 1), l = v.seq_len[r];
  const u8 *s = v.seqb + v.seq_off[r];
  u32 w[4] = {0, 0, 0, 0};
  // Fast path: all 16 amino acids belong to this element and their 48 nucleotides can be fetched as three 16-byte
  // windows (five aligned word loads + funnel shifts each) instead of 48 byte loads.
  const u32 i0 = start + 3u * j;  // first nucleotide index (on the translated strand)
  const u32 valid = plen[e] - j;  // amino acids of this chunk that exist (>= 1); the rest of the slot is padding
  // the 48-byte window may run past the end of the sequence (forward frames) or below its start (reverse frames):
  // it only has to stay inside the buffer
  const u64 sbase = (u64)v.seq_off[r];
  bool fast = f > 0 ? true : sbase + l >= 48ull + i0;
  u64 lo_addr = 0;
  if (fast) {
    lo_addr = f > 0 ? sbase + i0 : sbase + l - 48u - i0;  // lowest source byte of the 48
    fast = (lo_addr & ~15ull) + 64 <= seq_limit;         // the four 16-byte loads stay inside the readable buffer
  }
  if (fast) {
    u32 nt[12];
    {
      // 48 bytes at any alignment sit inside four aligned 16-byte loads (offset <= 15, 15 + 48 <= 64); the word and
      // byte offsets are resolved with selects and funnel shifts
      const uint4 *vp = reinterpret_cast<const uint4 *>(v.seqb + (lo_addr & ~15ull));
      const uint4 q0 = vp[0], q1 = vp[1], q2 = vp[2], q3 = vp[3];
      const u32 x[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
      const u32 off = (u32)(lo_addr & 15ull);
      const bool w1 = (off & 4u) != 0, w2 = (off & 8u) != 0;
      const u32 sh = (off & 3u) * 8u;
      u32 z[15], y[13];
#pragma unroll
      for (int q = 0; q < 15; q++) z[q] = w1 ? x[q + 1] : x[q];
#pragma unroll
      for (int q = 0; q < 13; q++) y[q] = w2 ? z[q + 2] : z[q];
#pragma unroll
      for (int q = 0; q < 12; q++) nt[q] = __funnelshift_r(y[q], y[q + 1], sh);
    }
    // branch-free inner loop: three code look-ups, one codon look-up, one OR into the output word per amino acid;
    // `special` collects the gap / not-a-nucleotide flags of the codons that exist
    u32 special = 0;
    if (f > 0) {
#pragma unroll
      for (int t = 0; t < (int)kTrAA; t++) {
        const u32 c0 = s_fwd[(nt[(3 * t) >> 2] >> (8 * ((3 * t) & 3))) & 0xffu];
        const u32 c1 = s_fwd[(nt[(3 * t + 1) >> 2] >> (8 * ((3 * t + 1) & 3))) & 0xffu];
        const u32 c2 = s_fwd[(nt[(3 * t + 2) >> 2] >> (8 * ((3 * t + 2) & 3))) & 0xffu];
        if ((u32)t < valid) special |= c0 | c1 | c2;
        const u32 x = s_lut[((c0 * 16u + c1) * 16u + c2) & 0xfffu];
        w[t >> 2] |= x << (8 * (t & 3));
      }
    } else {  // codon t reads the window from its top: bytes 47-3t, 46-3t, 45-3t
#pragma unroll
      for (int t = 0; t < (int)kTrAA; t++) {
        const u32 c0 = s_rev[(nt[(47 - 3 * t) >> 2] >> (8 * ((47 - 3 * t) & 3))) & 0xffu];
        const u32 c1 = s_rev[(nt[(46 - 3 * t) >> 2] >> (8 * ((46 - 3 * t) & 3))) & 0xffu];
        const u32 c2 = s_rev[(nt[(45 - 3 * t) >> 2] >> (8 * ((45 - 3 * t) & 3))) & 0xffu];
        if ((u32)t < valid) special |= c0 | c1 | c2;
        const u32 x = s_lut[((c0 * 16u + c1) * 16u + c2) & 0xfffu];
        w[t >> 2] |= x << (8 * (t & 3));
      }
    }
    if ((special & 0x30u) == 0) {
      if (c.init_m && j == 0 && (w[0] & 0x80u)) w[0] = (w[0] & ~0xffu) | (u32)'M';  // bit 7 of a table entry: start codon
#pragma unroll
      for (int q = 0; q < 4; q++) w[q] &= 0x7f7f7f7fu;
      if (c.clean) {
#pragma unroll
        for (int q = 0; q < 4; q++) {  // '*' (0x2a) -> 'X' (0x58): bytes are < 0x80, so the zero-byte test is exact
          const u32 y = w[q] ^ 0x2a2a2a2au;
          const u32 z = (((y + 0x7f7f7f7fu) | y) & 0x80808080u) ^ 0x80808080u;  // 0x80 where the byte was '*'
          w[q] ^= (z >> 7) * (u32)('*' ^ 'X');
        }
      }
      *reinterpret_cast<uint4 *>(prot + a0) = make_uint4(w[0], w[1], w[2], w[3]);
      return;
    }
    // a gap or a byte outside the alphabet among the chunk's codons: the careful loop below decides '-' / 'X' / error
  }
  // slow path (window would leave the buffer: first / last records of a block): byte loads, valid codons only
#pragma unroll 1
  for (u32 t = 0; t < valid && t < kTrAA; t++) {
    const u32 i = i0 + 3u * t;
    u32 c0, c1, c2;
    if (f > 0) {
      c0 = s_fwd[s[i]];
      c1 = s_fwd[s[i + 1]];
      c2 = s_fwd[s[i + 2]];
    } else {  // codon i of the reverse complement
      c0 = s_rev[s[l - 1 - i]];
      c1 = s_rev[s[l - 2 - i]];
      c2 = s_rev[s[l - 3 - i]];
    }
    u8 aa;
    bool init = false;
    if (c0 == 0x10 && c1 == 0x10 && c2 == 0x10) aa = '-';
    else if ((c0 | c1 | c2) & 0x30u) {
      aa = 'X';
      if (!c.allow_unknown) atomicMin((unsigned long long *)&st->err, ((unsigned long long)r << 4) | EK_UNKNOWN_CODON);
    } else {
      const u8 x = s_lut[(c0 << 8) | (c1 << 4) | c2];
      aa = x & 0x7f;
      init = (x & 0x80) != 0;
    }
    if (c.init_m && j + t == 0 && init) aa = 'M';
    if (c.clean && aa == '*') aa = 'X';
    prot[a0 + t] = aa;
  }
}

// --trim + the element views handed to the record formatter
__global__ void k_tr_views(RecViews v, TrCfg c, const u64 *__restrict__ poff, const u32 *__restrict__ plen,
                           const u8 *__restrict__ prot,
                           const u32 *__restrict__ hdr_off, const u32 *__restrict__ hdr_len, u32 *name_off, u32 *name_len,
                           u32 *seq_off, u32 *seq_len) {
  const u32 e = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 n_el = v.n_rec * c.nf;
  if (e >= n_el) return;
  const u32 r = e / c.nf;
  u32 pl = plen[e];  // poff[] holds the padded slots
  const u8 *p = prot + poff[e];
  if (c.trim)
    while (pl && (p[pl - 1] == 'X' || p[pl - 1] == '*')) pl--;
  seq_off[e] = (u32)poff[e];
  seq_len[e] = pl;
  if (hdr_off) {
    name_off[e] = hdr_off[e];
    name_len[e] = hdr_len[e];
  } else {
    name_off[e] = v.name_off[r];
    name_len[e] = v.name_len[r];
  }
}

__device__ __forceinline__ u32 frame_digits(int f) { return f < 0 ? 2u : 1u; }

// -F/--append-frame header: ID + "_frame=" + f + " " + Desc   (translate.go:134)
__global__ void k_tr_hdr_len(RecViews v, TrCfg c, const u32 *__restrict__ id_len, const u32 *__restrict__ desc_len,
                             u32 *__restrict__ hlen) {
  const u32 e = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 n_el = v.n_rec * c.nf;
  if (e > n_el) return;
  if (e == n_el) { hlen[e] = 0; return; }
  const u32 r = e / c.nf, fi = e - r * c.nf;
  hlen[e] = id_len[r] + 7u + frame_digits(c.frames[fi]) + 1u + desc_len[r];
}
__global__ void k_tr_hdr_fill(RecViews v, TrCfg c, const u32 *__restrict__ id_off, const u32 *__restrict__ id_len,
                              const u32 *__restrict__ desc_off, const u32 *__restrict__ desc_len,
                              const u32 *__restrict__ hoff, u8 *__restrict__ harena) {
  const u32 e = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 n_el = v.n_rec * c.nf;
  if (e >= n_el) return;
  const u32 r = e / c.nf, fi = e - r * c.nf;
  u8 *o = harena + hoff[e];
  const u8 *id = v.in + id_off[r];
  for (u32 i = 0; i < id_len[r]; i++) *o++ = id[i];
  const char *lit = "_frame=";
  for (int i = 0; i < 7; i++) *o++ = (u8)lit[i];
  const int f = c.frames[fi];
  if (f < 0) *o++ = '-';
  *o++ = (u8)('0' + (f < 0 ? -f : f));
  *o++ = ' ';
  const u8 *d = v.in + desc_off[r];
  for (u32 i = 0; i < desc_len[r]; i++) *o++ = d[i];
}

static u32 iupac_mask(u8 c) {  // bit0=T/U bit1=C bit2=A bit3=G (NCBI TCAG order)
  switch (c | 32) {
    case 't': case 'u': return 1; case 'c': return 2; case 'a': return 4; case 'g': return 8;
    case 'r': return 12; case 'y': return 3; case 's': return 10; case 'w': return 5;
    case 'k': return 9; case 'm': return 6; case 'b': return 11; case 'd': return 13;
    case 'h': return 7; case 'v': return 14; case 'n': return 15;
  }
  return 0;
}

// host tables: base -> 4-bit IUPAC code (16 = '-') for both strands, codon (3 codes) -> amino acid | 0x80 when every
// expansion is a start codon (512 + 4096 bytes)
void Engine::translate_host_tables(u8 *tab) {
  const GCode *gc = find_gcode(o_.TranslTable);
  const u8 *pair = alphabet_pair(alphabet_);
  for (int b = 0; b < 256; b++) {
    tab[b] = b == '-' ? 16 : (u8)iupac_mask((u8)b);
    const u8 pb = pair[b];
    tab[256 + b] = pb == '-' ? 16 : (u8)iupac_mask(pb);
  }
  u8 *lut = tab + 512;
  for (u32 m0 = 0; m0 < 16; m0++)
    for (u32 m1 = 0; m1 < 16; m1++)
      for (u32 m2 = 0; m2 < 16; m2++) {
        u8 out = 0;
        if (m0 && m1 && m2) {
          int aa = 0;
          bool init = true;
          for (int i = 0; i < 4; i++) if (m0 >> i & 1)
            for (int j = 0; j < 4; j++) if (m1 >> j & 1)
              for (int kk = 0; kk < 4; kk++) if (m2 >> kk & 1) {
                const int idx = i * 16 + j * 4 + kk;
                const int a = gc->aas[idx];
                if (gc->starts[idx] != 'M') init = false;
                if (aa == 0) aa = a;
                else if (aa != a) aa = 'X';
              }
          out = (u8)aa | (init ? 0x80 : 0);
        }
        lut[(m0 << 8) | (m1 << 4) | m2] = out;
      }
}

// translate on FASTA wrapped at one width (or unwrapped), sequences read in place: record table from one streaming
// pass (k_fasta_tile.cu), element sizes from the record table alone, one output-driven kernel (k_translate_tile.cu).
// kFusedFallback when the options or the block are outside that path (--trim needs the proteins before their sizes are
// known, -F builds new headers, FASTQ input, ragged wrapping, no final newline).
int Engine::op_translate_tile(const u8 *d_in, u32 n, BlockOut &bo) {
  if (n == 0 || !fused_ok_ || o_.Trim || o_.AppendFrame || o_.LineWidth > 1000 || o_.frames.empty() || o_.frames.size() > 8 || getenv("BSK_NO_TRANSLATE_TILE"))
    return kFusedFallback;
  bool fastq = false, ok = false;
  const int saved_alpha = alphabet_;
  const bool saved_known = alphabet_known_;
  auto decline = [&]() { alphabet_ = saved_alpha; alphabet_known_ = saved_known; return kFusedFallback; };
  if (!alphabet_known_ || first_block_) {
    part_width_known_ = false;
    int rc = first_record_alphabet(d_in, n, fastq, ok, true);
    if (rc != BSK_OK) return rc;
    if (!ok || fastq) return decline();
  } else if (part_fastq_) {
    return kFusedFallback;
  }
  const bool nucleic = alphabet_ == AB_DNA || alphabet_ == AB_DNARED || alphabet_ == AB_RNA || alphabet_ == AB_RNARED;
  if (!nucleic) return decline();  // the general path reports the error
  // line width of the block: the first sequence line (inside the probe) that is followed by another sequence line
  u32 width = part_width_;
  if (first_block_ || !part_width_known_) {
    width = 0;
    const u8 *p = h_probe_.as<u8>();
    const u32 pn = n < (256u << 10) ? n : (256u << 10);
    u32 ls = 0;
    bool prev_seq = false;
    u32 prev_len = 0;
    for (u32 i = 0; i <= pn && !width; i++) {
      if (i == pn || p[i] == '\n') {
        if (i == pn) break;
        const bool is_hdr = p[ls] == '>';
        if (!is_hdr && prev_seq) width = prev_len;
        prev_seq = !is_hdr;
        prev_len = i - ls;
        ls = i + 1;
      }
    }
    if (width > 1000) return decline();
    part_width_ = width;
    part_width_known_ = true;
  }
  if (!n_sm_) {
    cudaDeviceProp prop;
    BSK_CUDA(cudaGetDeviceProperties(&prop, device_ >= 0 ? device_ : 0));
    n_sm_ = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 1;
  }
  reset_status();
  const u32 n_tiles = k::fasta_tile_tiles(n);
  const size_t n_spans = (size_t)n_tiles * k::fasta_tile_spans_per_tile() + 4;
  k::FastaTileArgs ia;
  memset(&ia, 0, sizeof ia);
  ia.in = d_in;
  ia.n = n;
  ia.width = width;
  ia.clean = b_op2_.get<u8>(n_spans);
  ia.hdr_cap = (u64)n / 32 + 1024;
  ia.hdr_off = b_op1_.get<u64>((size_t)ia.hdr_cap * 2);
  ia.hdr_nl = ia.hdr_off + ia.hdr_cap;
  ia.tile_nl = b_tile_cnt_.get<u32>((size_t)n_tiles + 1);
  ia.st = d_status_;
  BSK_CUDA(cudaMemsetAsync(ia.clean, 0, n_spans, stream));
  BSK_CUDA(cudaMemsetAsync(ia.tile_nl + n_tiles, 0, 4, stream));
  k::fasta_index_tile(ia, n_sm_, stream);
  launches_++;
  fetch_status();
  const u64 n_hdr = h_status_->counters[6];
  if (h_status_->counters[0] || h_status_->counters[7] || n_hdr > ia.hdr_cap || n_hdr * o_.frames.size() >= 0xFFFFFFF0ull) return decline();
  n_rec_ = (u32)n_hdr;
  const size_t R = (size_t)n_rec_ + 1;
  u64 *hs_off = b_op5_.get<u64>(R * 2), *hs_nl = hs_off + R;
  prim::sort_pairs_u64_u64(ia.hdr_off, hs_off, ia.hdr_nl, hs_nl, n_rec_, 0, 32, b_tmp_, stream);
  u32 *tile_base = b_tile_base_.get<u32>((size_t)n_tiles + 1);
  prim::excl_scan_u32(ia.tile_nl, tile_base, (size_t)n_tiles + 1, b_tmp_, stream);
  u32 *rec = b_rec_.get<u32>(R * 5);
  u32 *name_off = rec, *name_len = rec + R, *seq_start = rec + 2 * R, *seq_nl = rec + 3 * R, *seq_len = rec + 4 * R;
  k::locate_records(d_in, n, hs_off, hs_nl, tile_base, n_tiles, n_rec_, name_off, name_len, seq_start, seq_nl, seq_len, stream);
  launches_++;
  BSK_CUDA(cudaEventRecord(ev_[1], stream));
  // the block state error reporting reads
  in_ = d_in;
  n_ = n;
  fastq_ = false;
  squeezed_ = false;
  if (first_block_) part_fastq_ = false;
  memset(&ra_, 0, sizeof ra_);
  ra_.head_off = name_off;
  ra_.head_len = name_len;
  ra_.seq_len = seq_len;
  views_ = RecViews{};
  views_.in = d_in;
  views_.seqb = d_in;
  views_.qualb = d_in;
  views_.name_off = name_off;
  views_.name_len = name_len;
  views_.seq_off = seq_start;
  views_.seq_len = seq_len;
  views_.qual_len = seq_len;
  views_.n_rec = n_rec_;
  bo.n_rec = n_rec_;
  if (n_rec_) any_record_ = true;
  if (n_rec_ == 0) { timings.fused_blocks++; return BSK_OK; }

  const u32 nf = (u32)o_.frames.size();
  const u32 n_el = n_rec_ * nf;
  int frames[8];
  for (u32 i = 0; i < 8; i++) frames[i] = i < nf ? o_.frames[i] : 1;
  const u32 wout = o_.LineWidth > 0 ? (u32)o_.LineWidth : 0;
  u32 *sizes = b_out_len_.get<u32>((size_t)n_el + 1);
  u64 *out_off = b_out_off_.get<u64>((size_t)n_el + 1);
  k::translate_sizes(name_len, seq_len, n_rec_, nf, frames, wout, sizes, d_status_, stream);
  launches_++;
  prim::excl_scan_u32_to_u64(sizes, out_off, (size_t)n_el + 1, b_tmp_, stream);
  u8 *hs = h_small_.as<u8>() + 8192;
  prim::copy_small(hs, out_off + n_el, 8, stream);
  // tables while the scan runs
  u8 *tab = h_small_.as<u8>();
  translate_host_tables(tab);
  const GCode *gc = find_gcode(o_.TranslTable);
  u8 *aa = tab + 512 + 4096;  // 64 + 64 entries for plain ACGT codons; base codes of the kernel: A 0, C 1, T 2, G 3
  u64 start_fwd = 0, start_rev = 0;
  static const int ncbi_of[4] = {2, 1, 0, 3};  // kernel code -> index in NCBI's TCAG order
  for (int idx = 0; idx < 64; idx++) {
    const int b0 = idx & 3, b1 = (idx >> 2) & 3, b2 = (idx >> 4) & 3;
    const int fi = ncbi_of[b0] * 16 + ncbi_of[b1] * 4 + ncbi_of[b2];
    const int ri = ncbi_of[b2 ^ 2] * 16 + ncbi_of[b1 ^ 2] * 4 + ncbi_of[b0 ^ 2];  // complement: code ^ 2
    u8 x = (u8)gc->aas[fi], y = (u8)gc->aas[ri];
    if (o_.Clean && x == '*') x = 'X';
    if (o_.Clean && y == '*') y = 'X';
    aa[idx] = x;
    aa[64 + idx] = y;
    if (gc->starts[fi] == 'M') start_fwd |= 1ull << idx;
    if (gc->starts[ri] == 'M') start_rev |= 1ull << idx;
  }
  u8 *d_tab = b_op3_.get<u8>(512 + 4096 + 128);
  BSK_CUDA(cudaMemcpyAsync(d_tab, tab, 512 + 4096 + 128, cudaMemcpyHostToDevice, stream));
  fetch_status();  // the sizes pass flags sequences shorter than one codon
  int rc = check_errors();
  if (rc != BSK_OK) return rc;
  u64 total;
  memcpy(&total, hs, 8);
  u8 *out = b_out_.get<u8>((size_t)total + 64);
  k::TranslateTileArgs ta;
  memset(&ta, 0, sizeof ta);
  ta.in = d_in;
  ta.n = n;
  ta.clean = ia.clean;
  ta.n_rec = n_rec_;
  ta.nf = nf;
  for (u32 i = 0; i < 8; i++) ta.frames[i] = frames[i];
  ta.name_off = name_off;
  ta.name_len = name_len;
  ta.seq_start = seq_start;
  ta.seq_len = seq_len;
  ta.width_in = width;
  ta.width_out = wout;
  ta.magic_in = width ? ((1ull << 40) + width - 1) / width : 0;
  ta.magic_out1 = ((1ull << 40) + wout) / (wout + 1ull);
  ta.tile_first = b_op4_.get<u32>((size_t)k::translate_tile_tiles(total) + 1);
  ta.out_off = out_off;
  ta.total = total;
  ta.code_fwd = d_tab;
  ta.code_rev = d_tab + 256;
  ta.lut = d_tab + 512;
  ta.aa_fwd = d_tab + 512 + 4096;
  ta.aa_rev = d_tab + 512 + 4096 + 64;
  ta.start_fwd = start_fwd;
  ta.start_rev = start_rev;
  ta.allow_unknown = o_.AllowUnknownCodon;
  ta.init_m = o_.InitCodonAsM;
  ta.clean_stop = o_.Clean;
  ta.out = out;
  ta.st = d_status_;
  main_begin();
  k::translate_tile(ta, stream);
  main_end();
  launches_ += 2;
  fetch_status();
  rc = check_errors();
  if (rc != BSK_OK) return rc;
  bo.d_data = out;
  bo.n = total;
  bo.n_elem = n_el;
  bo.d_elem_off = want_elem_off ? out_off : nullptr;
  timings.fused_blocks++;
  return BSK_OK;
}

int Engine::op_translate(BlockOut &bo) {
  if (n_rec_ == 0) return BSK_OK;
  // translate.go:116-122: the partition's alphabet must be DNA/RNA; a parse error on record 0 comes first
  const bool nucleic = alphabet_ == AB_DNA || alphabet_ == AB_DNARED || alphabet_ == AB_RNA || alphabet_ == AB_RNARED;
  if (!nucleic && first_block_) {
    if (h_status_->err != kNoErr && (h_status_->err >> 4) == 0) return check_errors();
    err = "command 'seqkit translate' only apply to DNA/RNA sequences";
    return BSK_ERR_DATA;
  }
  TrCfg c;
  memset(&c, 0, sizeof c);
  c.nf = (u32)std::min<size_t>(o_.frames.size(), 66);
  for (u32 i = 0; i < c.nf; i++) c.frames[i] = o_.frames[i];
  c.allow_unknown = o_.AllowUnknownCodon;
  c.init_m = o_.InitCodonAsM;
  c.clean = o_.Clean;
  c.trim = o_.Trim;
  if (c.nf == 0) return check_errors();
  const u64 n_el64 = (u64)n_rec_ * c.nf;
  if (n_el64 >= 0xFFFFFFF0ull) { err = "translate: too many (record, frame) elements in one block"; return BSK_ERR_DATA; }
  const u32 n_el = (u32)n_el64;

  u8 *tab = h_small_.as<u8>();
  translate_host_tables(tab);
  u8 *d_tab = b_op1_.get<u8>(512 + 4096);
  BSK_CUDA(cudaMemcpyAsync(d_tab, tab, 512 + 4096, cudaMemcpyHostToDevice, stream));

  u32 *plen = b_op2_.get<u32>((size_t)n_el + 1);
  u64 *poff = b_op3_.get<u64>((size_t)n_el + 1);
  u32 *ppad = b_op8_.get<u32>((size_t)n_el + 1);
  BSK_LAUNCH_FLAT(k_tr_len, (n_el + 1 + 255) / 256, 256, 0, stream, views_, c, plen, ppad, d_status_);
  launches_++;
  prim::excl_scan_u32_to_u64(ppad, poff, (size_t)n_el + 1, b_tmp_, stream);
  u8 *hs = h_small_.as<u8>() + 8192;
  prim::copy_small(hs, poff + n_el, 8, stream);
  fetch_status();  // also syncs the copy above; errors are judged after the codon pass (earliest record wins)
  int rc = BSK_OK;
  u64 ptotal;
  memcpy(&ptotal, hs, 8);
  if (ptotal >= 0xFFFFFFF0ull) { err = "translate: protein arena exceeds 4 GiB in one block"; return BSK_ERR_DATA; }
  u8 *prot = b_op4_.get<u8>((size_t)ptotal + 64);
  if (ptotal) {
    main_begin();
    BSK_LAUNCH(k_translate, (u32)((ptotal + 256ull * kTrAA - 1) / (256ull * kTrAA)), 256, 0, stream, views_, c, poff, ptotal,
               d_tab, d_tab + 256, d_tab + 512, prot, d_status_, views_.seqb == in_ ? (u64)n_ : ~0ull, plen);
    main_end();
    launches_++;
    fetch_status();
  }
  rc = check_errors();
  if (rc != BSK_OK) return rc;
  // headers
  const u32 *hdr_off = nullptr, *hdr_len = nullptr;
  const u8 *name_base = in_;
  if (o_.AppendFrame) {
    const size_t R = (size_t)n_rec_ + 1;
    u32 *ids = b_id_.get<u32>(R * 4);
    k::id_desc(views_, o_.IDNCBI ? 1 : 0, ids, ids + R, ids + 2 * R, ids + 3 * R, stream);
    u32 *hl = b_op5_.get<u32>(((size_t)n_el + 1) * 2);
    u32 *ho = hl + n_el + 1;
    BSK_LAUNCH_FLAT(k_tr_hdr_len, (n_el + 1 + 255) / 256, 256, 0, stream, views_, c, ids + R, ids + 3 * R, hl);
    prim::excl_scan_u32(hl, ho, (size_t)n_el + 1, b_tmp_, stream);
    prim::copy_small(hs, ho + n_el, 4, stream);
    BSK_CUDA(cudaStreamSynchronize(stream));
    u32 htotal;
    memcpy(&htotal, hs, 4);
    u8 *harena = b_op6_.get<u8>((size_t)htotal + 64);
    BSK_LAUNCH_FLAT(k_tr_hdr_fill, (n_el + 255) / 256, 256, 0, stream, views_, c, ids, ids + R, ids + 2 * R, ids + 3 * R, ho,
                    harena);
    launches_ += 3;
    hdr_off = ho;
    hdr_len = hl;
    name_base = harena;
  }
  u32 *ev = b_op7_.get<u32>(((size_t)n_el + 1) * 4);
  const size_t E = (size_t)n_el + 1;
  BSK_LAUNCH_FLAT(k_tr_views, (n_el + 255) / 256, 256, 0, stream, views_, c, poff, plen, prot, hdr_off, hdr_len, ev, ev + E,
                  ev + 2 * E, ev + 3 * E);
  launches_++;
  RecViews saved = views_;
  const u32 saved_rec = n_rec_;
  views_.in = name_base;
  views_.seqb = prot;
  views_.qualb = prot;
  views_.name_off = ev;
  views_.name_len = ev + E;
  views_.seq_off = ev + 2 * E;
  views_.seq_len = ev + 3 * E;
  views_.qual_off = ev + 2 * E;
  views_.qual_len = ev + 3 * E;
  views_.n_rec = n_el;
  n_rec_ = n_el;
  EmitCfg cfg;
  cfg.marker = '>';
  cfg.print_name = 1;
  cfg.print_seq = 1;
  cfg.print_qual = 0;
  cfg.plus_line = 0;
  cfg.reverse = 0;
  cfg.width = o_.LineWidth > 0 ? (u32)o_.LineWidth : 0;
  rc = emit_records(cfg, nullptr, nullptr, bo);
  views_ = saved;
  n_rec_ = saved_rec;
  return rc;
}

}  // namespace bsk
