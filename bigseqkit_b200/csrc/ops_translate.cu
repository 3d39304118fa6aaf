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

  // host tables: base -> 4-bit IUPAC code (16 = '-'), codon (3 codes) -> amino acid | 0x80 when every expansion is a start codon
  const GCode *gc = find_gcode(o_.TranslTable);
  u8 *tab = h_small_.as<u8>();
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
  u8 *d_tab = b_op1_.get<u8>(512 + 4096);
  BSK_CUDA(cudaMemcpyAsync(d_tab, tab, 512 + 4096, cudaMemcpyHostToDevice, stream));

  u32 *plen = b_op2_.get<u32>((size_t)n_el + 1);
  u64 *poff = b_op3_.get<u64>((size_t)n_el + 1);
  u32 *ppad = b_op8_.get<u32>((size_t)n_el + 1);
  BSK_LAUNCH_FLAT(k_tr_len, (n_el + 1 + 255) / 256, 256, 0, stream, views_, c, plen, ppad, d_status_);
  launches_++;
  prim::excl_scan_u32_to_u64(ppad, poff, (size_t)n_el + 1, b_tmp_, stream);
  u8 *hs = h_small_.as<u8>() + 8192;
  BSK_CUDA(cudaMemcpyAsync(hs, poff + n_el, 8, cudaMemcpyDeviceToHost, stream));
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
    BSK_CUDA(cudaMemcpyAsync(hs, ho + n_el, 4, cudaMemcpyDeviceToHost, stream));
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
