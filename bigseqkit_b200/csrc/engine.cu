// engine.cu -- common per-block pipeline: index -> parse -> (squeeze) -> alphabet -> operator.
#include "engine.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "op_state.h"
#include "prims.h"

namespace bsk {

// per-operator state that outlives a block (rmdup key table and history, locate / grep pattern tables)
void Engine::free_op_state() {
  rmdup_state_free(rm_);
  rm_ = nullptr;
  delete pats_;
  pats_ = nullptr;
}
void Engine::reset_op_state() { rmdup_state_reset(rm_); }

Engine::Engine(Op op, const Opts &o, int device) : op_(op), o_(o), device_(device) {
  if (device_ >= 0) BSK_CUDA(cudaSetDevice(device_));
  BSK_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  BSK_CUDA(cudaMalloc((void **)&d_status_, sizeof(DevStatus)));
  BSK_CUDA(cudaHostAlloc((void **)&h_status_, sizeof(DevStatus), cudaHostAllocMapped | cudaHostAllocPortable));
  h_small_.reserve(16384);
  for (auto &e : ev_) BSK_CUDA(cudaEventCreate(&e));
  // constant tables: class[256] valid[256] lut[256] gap[256] aux[256] qpow[256 doubles]
  b_tables_.reserve(256 * 5 + 256 * sizeof(double) + 64);
  u8 *base = b_tables_.as<u8>();
  t_class_ = base;
  t_valid_ = base + 256;
  t_lut_ = base + 512;
  t_gap_ = base + 768;
  t_aux_ = base + 1024;
  t_qpow_ = reinterpret_cast<double *>(base + 1280);
  alphabet_ = o_.alphabet;
  alphabet_known_ = false;
  fused_ok_ = getenv("BSK_NO_FUSED") == nullptr;      // debugging aid: force the general index/parse/emit path
  inplace_ok_ = getenv("BSK_NO_INPLACE") == nullptr;  // debugging aid: skip the same-layout FASTQ kernel
}

Engine::~Engine() {
  if (device_ >= 0) cudaSetDevice(device_);
  if (stream) cudaStreamSynchronize(stream);
  if (d_status_) cudaFree(d_status_);
  if (h_status_) cudaFreeHost(h_status_);
  for (auto &e : ev_) if (e) cudaEventDestroy(e);
  for (int i = 0; i < 2; i++) {
    if (ev_in_done_[i]) cudaEventDestroy(ev_in_done_[i]);
    if (ev_in_free_[i]) cudaEventDestroy(ev_in_free_[i]);
    if (ev_out_ready_[i]) cudaEventDestroy(ev_out_ready_[i]);
    if (ev_out_free_[i]) cudaEventDestroy(ev_out_free_[i]);
  }
  if (s_in_) { cudaStreamSynchronize(s_in_); cudaStreamDestroy(s_in_); }
  if (s_out_) { cudaStreamSynchronize(s_out_); cudaStreamDestroy(s_out_); }
  comm_free();
  free_op_state();
  if (stream) cudaStreamDestroy(stream);
}

int Engine::reset() {
  hist_.clear();
  q20_ = q30_ = gap_ = 0;
  stats_type_.clear();
  stats_type_set_ = false;
  keys_host_.clear();
  rmdup_removed = 0;
  grep_count = 0;
  alphabet_ = o_.alphabet;
  alphabet_known_ = false;
  first_block_ = true;
  any_record_ = false;
  range_seen_ = 0;
  reset_op_state();
  return BSK_OK;
}

int Engine::stage_device(const u8 *in, size_t n, void **d_ptr) {
  if (n >= kMaxBlockBytes) { err = "bsk_stage_device: partition must be smaller than 4 GiB - 1 MiB"; return BSK_ERR_ARG; }
  if (device_ >= 0) BSK_CUDA(cudaSetDevice(device_));
  u8 *d = b_in_.get<u8>(n + 64);
  if (n) BSK_CUDA(cudaMemcpyAsync(d, in, n, cudaMemcpyHostToDevice, stream));
  BSK_CUDA(cudaStreamSynchronize(stream));
  *d_ptr = d;
  return BSK_OK;
}

// The status words travel by kernel, not by copy engine (prims.h: copy_small): the pipeline of bsk_run_buffer keeps both
// engines busy with block-sized copies, and a small copy of this stream would wait behind them.
static __global__ void k_status_reset(DevStatus *st) {
  if (threadIdx.x == 0) {
    DevStatus z;
    memset(&z, 0, sizeof z);
    z.err = kNoErr;
    z.guess_mask = 0xffffffffu;
    *st = z;
  }
}
void Engine::reset_status() {
  memset(h_status_, 0, sizeof(DevStatus));
  h_status_->err = kNoErr;
  h_status_->guess_mask = 0xffffffffu;
  BSK_LAUNCH_FLAT(k_status_reset, 1, 32, 0, stream, d_status_);
}

void Engine::fetch_status() {
  prim::copy_small(h_status_, d_status_, (u32)sizeof(DevStatus), stream);
  BSK_CUDA(cudaStreamSynchronize(stream));
}

// host tables that do not depend on the guessed alphabet
void Engine::upload_tables() {
  u8 *h = h_small_.as<u8>();
  memset(h, 0, 1280 + 256 * sizeof(double));
  alphabet_class_masks(h);  // class
  for (unsigned char c : o_.GapLetters) h[768 + c] = 1;
  double *qp = reinterpret_cast<double *>(h + 1280);
  for (int q = 0; q < 256; q++) qp[q] = pow(10, (double)(q - o_.QualAsciiBase) / -10);  // Seq.AvgQual term
  for (int c = 0; c < 256; c++) h[512 + c] = (u8)c;
  BSK_CUDA(cudaMemcpyAsync(b_tables_.p, h, 1280 + 256 * sizeof(double), cudaMemcpyHostToDevice, stream));
  BSK_CUDA(cudaStreamSynchronize(stream));  // h_small_ is reused
}

void Engine::set_views_default() {
  views_.in = in_;
  views_.seqb = squeezed_ ? b_seq_arena_.as<u8>() : in_;
  views_.qualb = squeezed_ ? b_qual_arena_.as<u8>() : in_;
  views_.name_off = ra_.head_off;
  views_.name_len = ra_.head_len;
  views_.seq_off = squeezed_ ? b_seq_aoff_.as<u32>() : ra_.seq_off;
  views_.seq_len = ra_.seq_len;
  views_.qual_off = squeezed_ ? b_qual_aoff_.as<u32>() : ra_.qual_off;
  views_.qual_len = ra_.qual_len;
  views_.n_rec = n_rec_;
}

// 4-line FASTQ reads: record index + parse in ONE streaming pass (k_rmdup_tile.cu without the hashing) instead of
// count + fill + parse.  kFusedFallback when the block is outside that grammar (the general passes below then run).
int Engine::prepare_block_tile(const u8 *d_in, u32 n) {
  if (n == 0 || !fused_ok_ || getenv("BSK_NO_TILE_INDEX")) return kFusedFallback;
  u8 *hs = h_small_.as<u8>();
  prim::copy_small(hs, d_in, 1, stream);
  BSK_CUDA(cudaStreamSynchronize(stream));
  if (hs[0] != '@') return kFusedFallback;
  if (!n_sm_) {
    cudaDeviceProp prop;
    BSK_CUDA(cudaGetDeviceProperties(&prop, device_ >= 0 ? device_ : 0));
    n_sm_ = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 1;
  }
  reset_status();
  if (first_block_) upload_tables();
  const u32 n_tiles = k::rmdup_tile_tiles(n);
  void *slots = b_op3_.get<u8>((size_t)n_tiles * k::rmdup_tile_slot_stride() * k::rmdup_tile_slot_bytes());
  u32 *tile_cnt = b_tile_cnt_.get<u32>((size_t)n_tiles + 1);
  BSK_CUDA(cudaMemsetAsync(tile_cnt, 0, ((size_t)n_tiles + 1) * 4, stream));
  k::rmdup_tile(d_in, n, slots, tile_cnt, d_status_, -1, n_sm_, stream);
  launches_++;
  u64 *tile_base = b_tile_base_.get<u64>((size_t)n_tiles + 1);
  prim::excl_scan_u32_to_u64(tile_cnt, tile_base, (size_t)n_tiles + 1, b_tmp_, stream);
  prim::copy_small(hs, tile_base + n_tiles, 8, stream);
  fetch_status();  // synchronises the stream
  if (h_status_->counters[0]) return kFusedFallback;
  u64 nrec;
  memcpy(&nrec, hs, 8);
  in_ = d_in;
  n_ = n;
  n_rec_ = (u32)nrec;
  n_nl_ = n_lines_ = 4 * n_rec_;
  fastq_ = true;
  squeezed_ = false;
  if (first_block_) part_fastq_ = true;
  ix_ = RecIndex{d_in, n, nullptr, nullptr, n_lines_, n_rec_, 1};
  const size_t R = (size_t)n_rec_ + 1;
  u32 *rec = b_rec_.get<u32>(R * 10);
  ra_.head_off = rec;
  ra_.head_len = rec + R;
  ra_.seq_line0 = rec + 2 * R;
  ra_.seq_line1 = rec + 3 * R;
  ra_.seq_off = rec + 4 * R;
  ra_.seq_len = rec + 5 * R;
  ra_.qual_line0 = rec + 6 * R;
  ra_.qual_line1 = rec + 7 * R;
  ra_.qual_off = rec + 8 * R;
  ra_.qual_len = rec + 9 * R;
  BSK_CUDA(cudaMemsetAsync(ra_.seq_len + n_rec_, 0, 4, stream));
  BSK_CUDA(cudaMemsetAsync(ra_.qual_len + n_rec_, 0, 4, stream));
  k::rmdup_tile_compact(slots, tile_cnt, tile_base, n_tiles, nullptr, nullptr, ra_, nullptr, stream);
  launches_++;
  seq_space_ = qual_space_ = n;
  set_views_default();
  h_status_->counters[0] = 0;
  return BSK_OK;
}

// index + parse + squeeze.  Leaves errors (if any) in h_status_ for check_errors().
int Engine::prepare_block(const u8 *d_in, u32 n) {
  {
    const int trc = prepare_block_tile(d_in, n);
    if (trc != kFusedFallback) return trc;
  }
  in_ = d_in;
  n_ = n;
  n_nl_ = n_rec_ = n_lines_ = 0;
  squeezed_ = false;
  reset_status();
  if (first_block_) upload_tables();
  if (n == 0) {
    fastq_ = false;
    ix_ = RecIndex{d_in, 0, nullptr, nullptr, 0, 0, 0};
    memset(&ra_, 0, sizeof ra_);
    set_views_default();
    h_status_->err = kNoErr;
    return BSK_OK;
  }
  const u32 n_tiles = (n + k::kIndexTile - 1) / k::kIndexTile;
  u64 *tile_cnt = b_tile_.get<u64>(n_tiles + 1);
  u64 *tile_base = b_tile_scan_.get<u64>(n_tiles + 1);
  BSK_CUDA(cudaMemsetAsync(tile_cnt + n_tiles, 0, sizeof(u64), stream));
  k::index_count(d_in, n, tile_cnt, n_tiles, stream);
  launches_++;
  prim::excl_scan_u64(tile_cnt, tile_base, n_tiles + 1, b_tmp_, stream);
  u8 *hs = h_small_.as<u8>();
  prim::copy_small(hs, tile_base + n_tiles, 8, stream);
  prim::copy_small(hs + 8, d_in, 1, stream);
  prim::copy_small(hs + 9, d_in + (n - 1), 1, stream);
  BSK_CUDA(cudaStreamSynchronize(stream));
  u64 tot;
  memcpy(&tot, hs, 8);
  n_nl_ = (u32)tot;
  n_rec_ = (u32)(tot >> 32) + 1;
  fastq_ = hs[8] == '@';
  n_lines_ = hs[9] == '\n' ? n_nl_ : n_nl_ + 1;
  if (first_block_) part_fastq_ = fastq_;

  u32 *ls = b_ls_.get<u32>((size_t)n_nl_ + 2);
  u32 *rl = b_rl_.get<u32>((size_t)n_rec_ + 1);
  k::index_fill(d_in, n, tile_base, ls, rl, n_tiles, stream);
  k::index_finish(ls, rl, n, n_nl_, n_rec_, n_lines_, stream);
  launches_ += 2;
  ix_ = RecIndex{d_in, n, ls, rl, n_lines_, n_rec_, fastq_ ? 1 : 0};

  const size_t R = (size_t)n_rec_ + 1;
  u32 *rec = b_rec_.get<u32>(R * 10);
  ra_.head_off = rec;
  ra_.head_len = rec + R;
  ra_.seq_line0 = rec + 2 * R;
  ra_.seq_line1 = rec + 3 * R;
  ra_.seq_off = rec + 4 * R;
  ra_.seq_len = rec + 5 * R;
  ra_.qual_line0 = rec + 6 * R;
  ra_.qual_line1 = rec + 7 * R;
  ra_.qual_off = rec + 8 * R;
  ra_.qual_len = rec + 9 * R;
  // the scans below read one element past n_rec
  BSK_CUDA(cudaMemsetAsync(ra_.seq_len + n_rec_, 0, 4, stream));
  BSK_CUDA(cudaMemsetAsync(ra_.qual_len + n_rec_, 0, 4, stream));
  k::parse_records(ix_, ra_, d_status_, stream);
  launches_++;
  fetch_status();
  seq_space_ = qual_space_ = n;
  if (h_status_->multiline) {
    u32 *sa = b_seq_aoff_.get<u32>(R);
    u32 *qa = b_qual_aoff_.get<u32>(R);
    prim::excl_scan_u32(ra_.seq_len, sa, R, b_tmp_, stream);
    prim::excl_scan_u32(ra_.qual_len, qa, R, b_tmp_, stream);
    u8 *h2 = h_small_.as<u8>();
    prim::copy_small(h2, sa + n_rec_, 4, stream);
    prim::copy_small(h2 + 4, qa + n_rec_, 4, stream);
    const bool try_uniform = !fastq_ && getenv("BSK_NO_UNIFORM_SQUEEZE") == nullptr;
    if (try_uniform) {  // FASTA wrapped at a fixed width (every writer's output): no per-line work needed
      BSK_CUDA(cudaMemsetAsync(&d_status_->counters[7], 0, 8, stream));
      k::lines_uniform(ix_, &d_status_->counters[7], stream);
      launches_++;
      prim::copy_small(h2 + 8, &d_status_->counters[7], 8, stream);
    }
    BSK_CUDA(cudaStreamSynchronize(stream));
    u32 stot, qtot;
    memcpy(&stot, h2, 4);
    memcpy(&qtot, h2 + 4, 4);
    u64 ragged = 1;
    if (try_uniform) memcpy(&ragged, h2 + 8, 8);
    u8 *sar = b_seq_arena_.get<u8>((size_t)stot + 64);
    u8 *qar = b_qual_arena_.get<u8>((size_t)qtot + 64);
    if (ragged == 0) k::squeeze_uniform(ix_, ra_, sa, sar, stot, stream);
    else k::squeeze_lines(ix_, ra_, sa, qa, sar, qar, stream);
    launches_++;
    squeezed_ = true;
    seq_space_ = stot;
    qual_space_ = qtot;
  }
  set_views_default();
  return BSK_OK;
}

// alphabet of the partition: SeqType, else guessed from the first record of the
// partition (bigseqkit-lib/helper.go:286-291)
int Engine::resolve_alphabet() {
  if (alphabet_known_) return BSK_OK;
  if (n_rec_ == 0) return BSK_OK;  // nothing seen yet; stays unknown
  k::guess_alphabet(views_, t_class_, (u32)(o_.AlphabetGuessSeqLength > 0 ? o_.AlphabetGuessSeqLength : 0), d_status_,
                    stream);
  launches_++;
  const u64 keep_err = h_status_->err;
  fetch_status();
  (void)keep_err;
  first_guess_ = alphabet_from_mask(h_status_->guess_mask, h_status_->guess_len == 0);
  if (alphabet_ == AB_NIL) alphabet_ = first_guess_;
  alphabet_known_ = true;
  return BSK_OK;
}

std::string Engine::describe_error(u64 rec, u32 kind) {
  char buf[768];
  u32 meta[4] = {0, 0, 0, 0};  // head_off, head_len, seq_len, qual_len
  BSK_CUDA(cudaMemcpy(&meta[0], ra_.head_off + rec, 4, cudaMemcpyDeviceToHost));
  BSK_CUDA(cudaMemcpy(&meta[1], ra_.head_len + rec, 4, cudaMemcpyDeviceToHost));
  BSK_CUDA(cudaMemcpy(&meta[2], views_.seq_len + rec, 4, cudaMemcpyDeviceToHost));
  BSK_CUDA(cudaMemcpy(&meta[3], views_.qual_len + rec, 4, cudaMemcpyDeviceToHost));
  std::string head(meta[1], '\0');
  if (meta[1]) BSK_CUDA(cudaMemcpy(&head[0], in_ + meta[0], meta[1], cudaMemcpyDeviceToHost));
  switch (kind) {
    case EK_UNMATCHED:  // bigseqkit-lib/helper.go:308-311
      snprintf(buf, sizeof buf, "seq('%s'): unmatched length of sequence (%u) and quality (%u)", head.c_str(), meta[2],
               meta[3]);
      return buf;
    case EK_VALIDATE: {
      u32 so = 0;
      BSK_CUDA(cudaMemcpy(&so, views_.seq_off + rec, 4, cudaMemcpyDeviceToHost));
      u32 len = meta[2];
      if (o_.ValidateSeqLength > 0 && len > (u32)o_.ValidateSeqLength) len = (u32)o_.ValidateSeqLength;
      std::string s(len, '\0');
      if (len) BSK_CUDA(cudaMemcpy(&s[0], views_.seqb + so, len, cudaMemcpyDeviceToHost));
      const u8 *valid = alphabet_valid(alphabet_);
      char bad = '?';
      for (char c : s)
        if (!valid[(u8)c]) { bad = c; break; }
      snprintf(buf, sizeof buf, "seq: invalid %s letter: %c", alphabet_name(alphabet_), bad);
      return buf;
    }
    case EK_TOO_SHORT:
      return "seq: sequence too short to translate";
    case EK_UNKNOWN_CODON:
      return "seq: unknown codon";
  }
  return "data error";
}

int Engine::check_errors() {
  if (h_status_->err == kNoErr) return BSK_OK;
  const u64 rec = h_status_->err >> 4;
  const u32 kind = (u32)(h_status_->err & 0xf);
  err = describe_error(rec, kind);
  return BSK_ERR_DATA;
}

// ------------------------------------------------------------------ generic emit
int Engine::emit_records(const EmitCfg &cfg, const u8 *keep, const u8 *lut, BlockOut &bo) {
  const size_t R = (size_t)n_rec_ + 1;
  u32 *olen = b_out_len_.get<u32>(R);
  u64 *ooff = b_out_off_.get<u64>(R);
  k::out_len(views_, cfg, keep, olen, stream);
  launches_++;
  prim::excl_scan_u32_to_u64(olen, ooff, R, b_tmp_, stream);
  u8 *hs = h_small_.as<u8>();
  prim::copy_small(hs, ooff + n_rec_, 8, stream);
  u32 *d_nsel = &d_status_->n_sel;
  u64 *elem = nullptr;
  if (want_elem_off) {
    elem = b_elem_.get<u64>(R + 1);
    if (keep) {
      prim::select_flagged_u64(ooff, keep, elem, d_nsel, n_rec_, b_tmp_, stream);
      prim::copy_small(hs + 8, d_nsel, 4, stream);
    }
  } else if (keep) {
    // still need the element count: count kept records through the same select on a scratch
    elem = b_elem_.get<u64>(R + 1);
    prim::select_flagged_u64(ooff, keep, elem, d_nsel, n_rec_, b_tmp_, stream);
    prim::copy_small(hs + 8, d_nsel, 4, stream);
  }
  // records printed exactly as they stand in the input (rmdup, grep, filters on single-line records) are compacted
  // as byte ranges instead of being re-formatted byte by byte
  const bool try_contig = !squeezed_ && n_rec_ && cfg.marker && cfg.print_name && cfg.print_seq && !cfg.reverse && !lut &&
                          views_.seqb == in_ && (fastq_ ? (cfg.print_qual && cfg.plus_line && views_.qualb == in_) : !cfg.print_qual) &&
                          getenv("BSK_NO_CONTIG") == nullptr;
  if (try_contig && !contig_known_) {
    BSK_CUDA(cudaMemsetAsync(&d_status_->counters[7], 0, 8, stream));
    k::contig_check(views_, cfg, keep, fastq_ ? 1 : 0, n_, &d_status_->counters[7], stream);
    launches_++;
    prim::copy_small(hs + 16, &d_status_->counters[7], 8, stream);
  }
  BSK_CUDA(cudaStreamSynchronize(stream));
  u64 total;
  memcpy(&total, hs, 8);
  u32 nsel = n_rec_;
  if (keep) memcpy(&nsel, hs + 8, 4);
  u64 not_contig = 1;
  if (try_contig && !contig_known_) memcpy(&not_contig, hs + 16, 8);
  else if (try_contig) not_contig = 0;
  u8 *out = b_out_.get<u8>((size_t)total + 64);
  main_begin();
  if (not_contig == 0) {
    k::emit_contig(views_, ooff, out, total, n_, stream);
    if (total) BSK_CUDA(cudaMemsetAsync(out + total - 1, '\n', 1, stream));  // input without a final newline
  } else {
    // 16-byte windows may over-read a source by up to 19 bytes: fine inside the ctx's own buffers (all carry 64
    // bytes of slack), not inside the caller's input
    const u64 no_limit = ~0ull;
    k::emit(views_, cfg, ooff, out, total, lut, views_.in == in_ ? (u64)n_ : no_limit,
            views_.seqb == in_ ? (u64)n_ : no_limit, views_.qualb == in_ ? (u64)n_ : no_limit, stream);
  }
  main_end();
  launches_++;
  bo.d_data = out;
  bo.n = total;
  bo.n_elem = nsel;
  bo.d_elem_off = nullptr;
  if (want_elem_off) {
    if (keep) {
      BSK_CUDA(cudaMemcpyAsync(elem + nsel, ooff + n_rec_, 8, cudaMemcpyDeviceToDevice, stream));
      bo.d_elem_off = elem;
    } else {
      bo.d_elem_off = ooff;
    }
  }
  return BSK_OK;
}

// ------------------------------------------------------------------ block dispatch
// the FIRST bracketed kernel of a block is the one that is timed (the operator's streaming pass; formatters that
// run behind it in the same block are not)
void Engine::main_begin() {
  if (main_timed_) return;
  BSK_CUDA(cudaEventRecord(ev_[2], stream));
}
void Engine::main_end() {
  if (main_timed_) return;
  BSK_CUDA(cudaEventRecord(ev_[3], stream));
  main_timed_ = true;
  timings.main_launches++;
}
// call after the stream has been synchronised
void Engine::accumulate_timings() {
  float a = 0, b = 0, c = 0;
  cudaEventElapsedTime(&a, ev_[0], ev_[1]);
  cudaEventElapsedTime(&b, ev_[1], ev_[4]);
  if (main_timed_) cudaEventElapsedTime(&c, ev_[2], ev_[3]);
  timings.index_ms += a;
  timings.op_ms += b;
  timings.main_ms += c;
  timings.total_ms += a + b;
}

int Engine::process_block(const u8 *d_in, u32 n, int64_t pid, BlockOut &bo) {
  bo = BlockOut();
  main_timed_ = false;
  BSK_CUDA(cudaEventRecord(ev_[0], stream));
  if (seq_fused_eligible()) {
    const int frc = op_seq_fused(d_in, n, bo);
    if (frc != kFusedFallback) {
      first_block_ = false;
      if (frc == BSK_OK) {
        BSK_CUDA(cudaEventRecord(ev_[4], stream));
        BSK_CUDA(cudaStreamSynchronize(stream));
        accumulate_timings();
      }
      return frc;
    }
    bo = BlockOut();
    BSK_CUDA(cudaEventRecord(ev_[0], stream));
  }
  if (op_ == OP_STATS) {
    const int frc = op_stats_tile(d_in, n, bo);
    if (frc != kFusedFallback) {
      first_block_ = false;
      if (frc == BSK_OK) {
        BSK_CUDA(cudaEventRecord(ev_[4], stream));
        BSK_CUDA(cudaStreamSynchronize(stream));
        accumulate_timings();
      }
      return frc;
    }
    bo = BlockOut();
    BSK_CUDA(cudaEventRecord(ev_[0], stream));
  }
  if (op_ == OP_LOCATE || op_ == OP_TRANSLATE) {
    const int frc = op_ == OP_LOCATE ? op_locate_tile(d_in, n, pid, bo) : op_translate_tile(d_in, n, bo);
    if (frc != kFusedFallback) {
      first_block_ = false;
      if (frc == BSK_OK) {
        BSK_CUDA(cudaEventRecord(ev_[4], stream));
        BSK_CUDA(cudaStreamSynchronize(stream));
        accumulate_timings();
      }
      return frc;
    }
    bo = BlockOut();
    BSK_CUDA(cudaEventRecord(ev_[0], stream));
  }
  if (op_ == OP_RMDUP || op_ == OP_RMDUP_PREPARE) {
    const int frc = op_rmdup_tile(d_in, n, bo, op_ == OP_RMDUP_PREPARE);
    if (frc != kFusedFallback) {
      first_block_ = false;
      if (frc == BSK_OK) {
        BSK_CUDA(cudaEventRecord(ev_[4], stream));
        BSK_CUDA(cudaStreamSynchronize(stream));
        accumulate_timings();
      }
      return frc;
    }
    bo = BlockOut();
    BSK_CUDA(cudaEventRecord(ev_[0], stream));
  }
  int rc = prepare_block(d_in, n);
  if (rc != BSK_OK) return rc;
  bo.n_rec = n_rec_;
  if (n_rec_) any_record_ = true;
  rc = resolve_alphabet();
  if (rc != BSK_OK) return rc;
  BSK_CUDA(cudaEventRecord(ev_[1], stream));
  switch (op_) {
    case OP_SEQ: rc = op_seq(bo); break;
    case OP_STATS: rc = op_stats(bo); break;
    case OP_RMDUP: rc = op_rmdup(bo, false); break;
    case OP_RMDUP_PREPARE: rc = op_rmdup(bo, true); break;
    case OP_TRANSLATE: rc = op_translate(bo); break;
    case OP_LOCATE: rc = op_locate(bo, pid); break;
    case OP_GREP: rc = op_grep(bo); break;
    case OP_SUBSEQ: rc = op_subseq(bo); break;
    case OP_FQ2FA: rc = op_fq2fa(bo); break;
    case OP_DUPLICATE: rc = op_duplicate(bo); break;
    case OP_RANGE: rc = op_range(bo); break;
    default: err = "unknown operator"; rc = BSK_ERR_ARG;
  }
  first_block_ = false;
  if (rc == BSK_OK) {
    BSK_CUDA(cudaEventRecord(ev_[4], stream));
    BSK_CUDA(cudaStreamSynchronize(stream));
    accumulate_timings();
  }
  return rc;
}

int Engine::run_device(const void *d_in, size_t n, int64_t pid, bsk_out *out) {
  memset(out, 0, sizeof *out);
  if (n >= kMaxBlockBytes) { err = "bsk_run_device: partition must be smaller than 4 GiB - 1 MiB"; return BSK_ERR_ARG; }
  if (((uintptr_t)d_in & 15) != 0) { err = "bsk_run_device: device pointer must be 16-byte aligned"; return BSK_ERR_ARG; }
  if (device_ >= 0) BSK_CUDA(cudaSetDevice(device_));
  launches_ = 0;
  timings = bsk_timings{};
  // a device call is one whole partition
  alphabet_ = o_.alphabet;
  alphabet_known_ = false;
  first_block_ = true;
  BlockOut bo;
  int rc = process_block(static_cast<const u8 *>(d_in), (u32)n, pid, bo);
  if (rc == BSK_OK && op_ == OP_GREP && o_.Count) rc = finish_grep_count(bo);
  BSK_CUDA(cudaStreamSynchronize(stream));
  timings.kernel_launches = launches_;
  timings.in_bytes = n;
  timings.out_bytes = bo.n;
  if (rc != BSK_OK) return rc;
  out->data = bo.d_data;
  out->n = bo.n;
  out->elem_off = bo.d_elem_off;
  out->n_elem = bo.n_elem;
  out->n_records = bo.n_rec;
  return BSK_OK;
}

}  // namespace bsk
