// ops_tile.cu -- host side of the single-pass tile kernels for short records: `seq` on 4-line FASTQ in same-layout
// mode (k_fastq_inplace.cu) and `stats` (k_stats_tile.cu); everything they decline goes to the general path.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "engine.h"
#include "prims.h"

namespace bsk {

// print mode + byte map of SeqTransform.Call (bigseqkit-lib/seq.go:151-163, 188-239), shared by both paths
void Engine::seq_emit_cfg(bool fastq, EmitCfg &cfg, u8 *lut, bool &need_lut) const {
  bool print_name = true, print_seq = true, print_qual = fastq;
  if (o_.Name && o_.Seq) {
  } else if (o_.Name) {
    print_seq = false;
    print_qual = false;
  } else if (o_.Seq) {
    print_name = false;
    print_qual = false;
  } else if (o_.Qual) {
    print_name = false;
    print_seq = false;
    print_qual = true;
  }
  cfg.marker = (print_name && print_seq) ? (fastq ? '@' : '>') : 0;
  cfg.print_name = print_name;
  cfg.print_seq = print_seq;
  cfg.print_qual = print_qual;
  cfg.plus_line = print_qual && !o_.Qual;
  cfg.reverse = o_.Reverse;
  int width = o_.LineWidth;
  if (o_.Seq || o_.Qual) width = 0;  // seq.go:106-108
  if (fastq) width = 0;              // seq.go:123
  cfg.width = width > 0 ? (u32)width : 0;
  // complement -> dna2rna / rna2dna -> case, one table
  need_lut = false;
  const u8 *pair = alphabet_pair(alphabet_ == AB_NIL ? AB_UNLIMIT : alphabet_);
  const bool comp = o_.Complement && alphabet_ != AB_UNLIMIT && alphabet_ != AB_NIL;
  const bool is_rna = alphabet_ == AB_RNA || alphabet_ == AB_RNARED, is_dna = alphabet_ == AB_DNA || alphabet_ == AB_DNARED;
  for (int c = 0; c < 256; c++) {
    u8 x = (u8)c;
    if (comp) x = pair[x];
    if (o_.Dna2rna && !is_rna) x = x == 't' ? 'u' : (x == 'T' ? 'U' : x);
    if (o_.Rna2dna && !is_dna) x = x == 'u' ? 't' : (x == 'U' ? 'T' : x);
    if (o_.LowerCase) { if (x >= 'A' && x <= 'Z') x = (u8)(x + 32); }
    else if (o_.UpperCase) { if (x >= 'a' && x <= 'z') x = (u8)(x - 32); }
    lut[c] = x;
    if (x != (u8)c) need_lut = true;
  }
}

bool Engine::seq_fused_eligible() const {
  if (!fused_ok_ || op_ != OP_SEQ) return false;
  if (o_.RemoveGaps || o_.MinQual > 0 || o_.MaxQual > 0) return false;
  if (o_.ValidateSeq || !(o_.alphabet == AB_NIL || o_.alphabet == AB_UNLIMIT)) return false;  // validation needs the general path
  if (o_.OnlyId && o_.IDNCBI) return false;
  return true;
}

// Alphabet of the partition from its first record (SeqParser.Read on record 0 + GuessAlphabetLessConservatively,
// bigseqkit-lib/helper.go:219-291), done on the host from a 256 KiB probe of the block.
int Engine::first_record_alphabet(const u8 *d_in, u32 n, bool &fastq, bool &ok, bool long_ok) {
  ok = false;
  const u32 probe = n < (256u << 10) ? n : (256u << 10);
  h_probe_.reserve(probe + 16);
  u8 *d = h_probe_.as<u8>();
  BSK_CUDA(cudaMemcpyAsync(d, d_in, probe, cudaMemcpyDeviceToHost, stream));
  BSK_CUDA(cudaStreamSynchronize(stream));
  fastq = d[0] == '@';
  const u8 marker = fastq ? '@' : '>';
  // end of record 0 (exclusive): next record start, or EOF when the probe covers the whole block
  u32 end = probe;
  bool found = probe == n;
  for (u32 L = 1; L < probe; L++) {
    if (d[L - 1] != '\n' || d[L] != marker) continue;
    if (fastq && L >= 3 && d[L - 3] == '\n' && d[L - 2] == '+') continue;
    end = L;
    found = true;
    break;
  }
  // first record longer than the probe: not a short-record block; with long_ok the guess is still made when the
  // probe holds as many bases as the guess looks at (contigs)
  if (!found && !long_ok) return BSK_OK;
  if (end > 0 && d[end - 1] == '\n') end--;  // ReadFixer strips one trailing newline
  u32 p = 0;
  while (p < end && d[p] != '\n') p++;  // header line
  std::string seq;
  const u32 limit = o_.AlphabetGuessSeqLength > 0 ? (u32)o_.AlphabetGuessSeqLength : 0xffffffffu;
  if (p < end) {
    p++;
    bool in_qual = false;
    while (p <= end && !in_qual) {
      u32 q = p;
      while (q < end && d[q] != '\n') q++;
      const bool terminated = q < end;
      if (fastq) {
        if (!terminated) break;  // unterminated segment in sequence mode is dropped
        if (q > p && d[p] == '+') { in_qual = true; break; }
      }
      if (seq.size() < limit) seq.append(reinterpret_cast<const char *>(d + p), q - p);
      if (!terminated) break;
      p = q + 1;
    }
  }
  if (!found && seq.size() < limit) return BSK_OK;
  first_seq_len_ = (u32)seq.size();
  first_rec_bytes_ = end + 1;
  if (seq.size() > limit) seq.resize(limit);
  u8 cm[256];
  alphabet_class_masks(cm);
  unsigned m = 0xffffffffu;
  for (unsigned char c : seq) m &= cm[c];
  first_guess_ = alphabet_from_mask(m, seq.empty());
  if (alphabet_ == AB_NIL) alphabet_ = first_guess_;
  alphabet_known_ = true;
  ok = true;
  return BSK_OK;
}

int Engine::op_seq_fused(const u8 *d_in, u32 n, BlockOut &bo) {
  if (n == 0) return kFusedFallback;
  bool fastq = false, ok = false;
  const int saved_alpha = alphabet_;
  const bool saved_known = alphabet_known_;
  if (!alphabet_known_ || first_block_) {
    int rc = first_record_alphabet(d_in, n, fastq, ok);
    if (rc != BSK_OK) return rc;
    if (!ok) { alphabet_ = saved_alpha; alphabet_known_ = saved_known; return kFusedFallback; }
  } else {
    fastq = part_fastq_;  // every block of a partition starts on a record of the partition's format
  }
  EmitCfg cfg;
  bool need_lut = false;
  u8 *hl = h_small_.as<u8>() + 1024;
  seq_emit_cfg(fastq, cfg, hl, need_lut);
  if (!fastq && cfg.print_qual) { alphabet_ = saved_alpha; alphabet_known_ = saved_known; return kFusedFallback; }  // -q on FASTA: error path
  reset_status();
  BSK_CUDA(cudaMemcpyAsync(t_lut_, hl, 256, cudaMemcpyHostToDevice, stream));
  BSK_CUDA(cudaEventRecord(ev_[1], stream));
  // FASTQ, whole record printed, nothing dropped or renamed: the output record has the layout of the input record
  if (inplace_ok_ && fastq && cfg.print_name && cfg.print_seq && cfg.print_qual && !o_.OnlyId && o_.MinLen <= 0 && o_.MaxLen <= 0) {
    const int irc = op_seq_inplace(d_in, n, fastq, cfg, hl, need_lut, bo);
    if (irc != kFusedFallback) {
      if (irc != BSK_OK) { alphabet_ = saved_alpha; alphabet_known_ = saved_known; }
      return irc;
    }
    reset_status();
  }
  // Everything else goes to the general path (tile index + window-based formatter).
  alphabet_ = saved_alpha;
  alphabet_known_ = saved_known;
  return kFusedFallback;
}

// `stats` on short records: one streaming kernel (k_stats_tile.cu), histogram + counters merged on the host.
// bigseqkit-lib/stats.go:48-117; the totals keep sum semantics.
int Engine::op_stats_tile(const u8 *d_in, u32 n, BlockOut &bo) {
  if (n == 0 || !fused_ok_ || getenv("BSK_NO_STATS_TILE")) return kFusedFallback;
  u8 gaps[4];
  int n_gap = 0;
  if (o_.All) {  // the kernel compares against up to four distinct gap letters
    for (unsigned char c : o_.GapLetters) {
      bool seen = false;
      for (int i = 0; i < n_gap; i++) seen = seen || gaps[i] == c;
      if (seen) continue;
      if (n_gap == 4) return kFusedFallback;
      gaps[n_gap++] = c;
    }
  }
  bool fastq = false, ok = false;
  const int saved_alpha = alphabet_;
  const bool saved_known = alphabet_known_;
  if (!alphabet_known_ || first_block_) {
    int rc = first_record_alphabet(d_in, n, fastq, ok);
    if (rc != BSK_OK) return rc;
    if (!ok) { alphabet_ = saved_alpha; alphabet_known_ = saved_known; return kFusedFallback; }
  } else {
    fastq = part_fastq_;
  }
  if (!n_sm_) {
    cudaDeviceProp prop;
    BSK_CUDA(cudaGetDeviceProperties(&prop, device_ >= 0 ? device_ : 0));
    n_sm_ = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 1;
  }
  reset_status();
  BSK_CUDA(cudaEventRecord(ev_[1], stream));
  const u32 bins = k::stats_tile_bins();
  u64 *d_hist = b_op1_.get<u64>(bins);
  BSK_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)bins * 8, stream));
  main_begin();
  k::stats_tile(d_in, n, d_hist, d_status_, fastq ? 1 : 0, o_.All ? 1 : 0, o_.fq_offset, gaps, n_gap,
                2u * first_rec_bytes_ + 64u, n_sm_, stream);
  main_end();
  launches_++;
  h_probe_.reserve((size_t)bins * 8 + 16);
  u64 *h_hist = h_probe_.as<u64>();
  BSK_CUDA(cudaMemcpyAsync(h_hist, d_hist, (size_t)bins * 8, cudaMemcpyDeviceToHost, stream));
  fetch_status();  // synchronises the stream
  if (h_status_->counters[0]) {
    alphabet_ = saved_alpha;
    alphabet_known_ = saved_known;
    main_timed_ = false;
    timings.main_launches--;
    return kFusedFallback;
  }
  for (u32 i = 0; i < bins; i++)
    if (h_hist[i]) hist_[i] += h_hist[i];
  if (o_.All) {
    q20_ += h_status_->counters[1];
    q30_ += h_status_->counters[2];
    gap_ += h_status_->counters[3];
  }
  const u64 nrec = h_status_->counters[5];
  if (!stats_type_set_) {  // stats.go:106-114 + bigseqkit/stats.go:109-130
    if (alphabet_ == AB_DNARED) stats_type_ = "DNA";
    else if (alphabet_ == AB_RNARED) stats_type_ = "RNA";
    else stats_type_ = alphabet_name(first_guess_);
    stats_type_set_ = true;
  }
  fastq_ = fastq;
  if (first_block_) part_fastq_ = fastq;
  n_rec_ = (u32)nrec;
  bo.n_rec = nrec;
  if (nrec) any_record_ = true;
  timings.fused_blocks++;
  return BSK_OK;
}

// `seq` on FASTQ in same-layout mode: one streaming kernel + the element-offset expansion.
int Engine::op_seq_inplace(const u8 *d_in, u32 n, bool fastq, const EmitCfg &cfg, const u8 *h_lut, bool need_lut, BlockOut &bo) {
  (void)h_lut;
  if (!n_sm_) {
    cudaDeviceProp prop;
    int dev = device_ >= 0 ? device_ : 0;
    BSK_CUDA(cudaGetDeviceProperties(&prop, dev));
    n_sm_ = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 1;
  }
  const u32 n_tiles = k::fastq_inplace_tiles(n);
  u8 *out = b_out_.get<u8>((size_t)n + 64);
  u32 *tile_cnt = b_tile_cnt_.get<u32>((size_t)n_tiles + 1);
  u16 *slots = b_slots_.get<u16>((size_t)n_tiles * k::fastq_inplace_slot_stride());
  BSK_CUDA(cudaMemsetAsync(tile_cnt, 0, ((size_t)n_tiles + 1) * 4, stream));
  BSK_CUDA(cudaMemsetAsync(&d_status_->counters[4], 0xff, 8, stream));
  main_begin();
  // lanes per record: 8 lanes x 8 words hold segments up to 250 B (reads), 32 x 4 up to ~500 B; longer ones take
  // the byte-pair path inside the kernel
  int group = first_seq_len_ <= 250 ? 8 : 32;
  if (const char *e = getenv("BSK_FQ_GROUP")) group = atoi(e) == 8 ? 8 : (atoi(e) == 4 ? 4 : 32);  // test hook: both lane groupings on any input
  // the newline scan first covers the halo as far as two records like the first one reach
  const u32 scan_halo = 2u * first_rec_bytes_ + 64u;
  k::fastq_inplace(d_in, n, out, t_lut_, tile_cnt, slots, d_status_, cfg.reverse ? 1 : 0, need_lut ? 1 : 0, group,
                   first_seq_len_, scan_halo, n_sm_, stream);
  main_end();
  launches_++;
  u64 *tile_base = b_tile_base_.get<u64>((size_t)n_tiles + 1);
  prim::excl_scan_u32_to_u64(tile_cnt, tile_base, (size_t)n_tiles + 1, b_tmp_, stream);
  // element offsets are expanded before the host knows the record count (one host round trip instead of two): the
  // array is sized from the partition's first record, with a re-run in the rare case that was too small
  u64 *elem = nullptr;
  u64 cap = 0;
  if (want_elem_off) {
    cap = (u64)n / (first_rec_bytes_ > 16 ? first_rec_bytes_ / 2 : 8) + 1024;
    if (b_elem_.cap / 8 > cap) cap = b_elem_.cap / 8;
    elem = b_elem_.get<u64>((size_t)cap);
    k::fastq_elem_expand(tile_cnt, tile_base, slots, elem, n_tiles, cap, d_status_, stream);
    launches_++;
  }
  u8 *hs = h_small_.as<u8>();
  prim::copy_small(hs, tile_base + n_tiles, 8, stream);
  fetch_status();  // synchronises the stream
  if (h_status_->counters[0]) {
    if (getenv("BSK_DEBUG")) {
      const u64 info = h_status_->counters[4];
      fprintf(stderr, "bsk: same-layout FASTQ kernel declined %llu tile(s); first: tile %u bad=%u rescan=%u n_own=%u n_lines=%u\n",
              (unsigned long long)h_status_->counters[0], (unsigned)(info >> 32), (unsigned)((info >> 31) & 1),
              (unsigned)((info >> 30) & 1), (unsigned)((info >> 16) & 0x3ff), (unsigned)(info & 0xffff));
    }
    main_timed_ = false;
    timings.main_launches--;
    return kFusedFallback;
  }
  u64 nrec;
  memcpy(&nrec, hs, 8);
  const u64 total = h_status_->counters[1];
  if (want_elem_off && nrec + 1 > cap) {  // more records than the first one suggested
    cap = nrec + 2;
    elem = b_elem_.get<u64>((size_t)cap);
    k::fastq_elem_expand(tile_cnt, tile_base, slots, elem, n_tiles, cap, d_status_, stream);
    launches_++;
  }
  fastq_ = fastq;
  if (first_block_) part_fastq_ = fastq;
  n_rec_ = (u32)nrec;
  bo.d_data = out;
  bo.n = total;
  bo.d_elem_off = elem;
  bo.n_elem = nrec;
  bo.n_rec = nrec;
  if (nrec) any_record_ = true;
  timings.fused_blocks++;
  return BSK_OK;
}

}  // namespace bsk
