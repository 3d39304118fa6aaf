// engine.h -- host-side operator objects behind the C ABI (include/bsk.h).
//
// One Engine == one reference operator instance between Before() and After()
// (bigseqkit-lib/seq.go:21-26 SeqTransform, stats.go Stats, rmdup.go RmDupPrepare/
// RmDupCheck, translate.go Translate, locate.go Locate, grep.go Grep, subseq.go
// SubseqTransform), bound to one CUDA device and one stream.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "bsk.h"
#include "devbuf.h"
#include "kernels.h"
#include "opts.h"

namespace bsk {

static const size_t kMaxBlockBytes = 0xFFF00000ull;  // offsets inside a block are u32

struct BlockOut {
  u8 *d_data = nullptr;
  u64 n = 0;
  u64 *d_elem_off = nullptr;  // n_elem + 1 entries (device) when wanted
  u64 n_elem = 0;
  u64 n_rec = 0;
};

class Engine {
 public:
  Engine(Op op, const Opts &o, int device);
  ~Engine();

  int run_device(const void *d_in, size_t n, int64_t pid, bsk_out *out);
  int run_buffer(const u8 *in, size_t n, int64_t pid, bsk_out *out);
  int reset();
  // partitions of one Union (bigseqkit-cli/helper.go:131-138): the rmdup key table / history and the Range index keep
  // running from call to call until reset()
  void set_union(bool on) { union_ = on; }
  int stage_device(const u8 *in, size_t n, void **d_ptr);  // host partition -> the ctx's own device buffer
  // file range -> operator -> file through a bounded ring of pinned slots (run_file.cu); out_fd < 0: nothing is written
  int run_stream(int fd, u64 off, u64 len, int64_t pid, int out_fd, u64 out_off, u64 *out_bytes, u64 *n_records, u64 *n_elem);

  // stats
  int stats_result(bsk_stats *out);
  int stats_add(const u64 *len, const u64 *cnt, size_t n, u64 q20, u64 q30, u64 gap, const char *type);
  int stats_merge_from(const Engine &src);
  int stats_dense_device(void *d_hist, size_t nbins, u64 *n_overflow);
  long stats_render(const char *file, const char *format, char *buf, size_t cap);

  // rmdup
  int rmdup_keys(const int64_t **keys, size_t *n);
  int rmdup_prepare_device(const void *d_in, size_t n, void *d_fp, size_t fp_cap, u64 *n_records);
  int rmdup_resolve_device(const void *d_all_fp, u64 n_before, bsk_out *out);
  int rmdup_dup_seqs(const char **data, size_t *n);
  int rmdup_dup_num(const char **data, size_t *n);

  // exchange steps (comm.cu): NCCL communicator bound to the ctx, or several ctxs of one process
  struct CommState;
  int comm_init(const uint8_t *id, int n_ranks, int rank);
  void comm_free();
  int comm_rank(int *rank, int *n_ranks) const;
  int output_offsets(u64 n_local, u64 *offset, u64 *total);
  int stats_allreduce();
  void stats_clear_totals() { hist_.clear(); q20_ = q30_ = gap_ = 0; stats_type_.clear(); stats_type_set_ = false; }
  int rmdup_sharded(const void *d_in, size_t n, bsk_out *out);
  int rmdup_prepare_local(const void *d_in, size_t n, u64 *n_records);  // index + hash, keys / second hashes stay on the device
  void rmdup_export_fp(u64 *d_fp);                                      // interleaved {key, second hash} of the prepared shard
  void rmdup_export_own_fp();
  int rmdup_resolve_after(Engine **e, int self, const u64 *counts, bsk_out *out);

  struct RmdupState;   // ops_rmdup.cu: key table + fingerprint history of the running partition
  struct PatternSet;   // ops_match.cu: needles of locate / grep on the device

  std::string err;
  int want_elem_off = 1;
  bsk_timings timings{};
  cudaStream_t stream = nullptr;
  u64 rmdup_removed = 0;
  u64 grep_count = 0;

 private:
  Op op_;
  Opts o_;
  int device_;

  // ---- per-partition state (alphabet is guessed on the first record of a partition)
  int alphabet_ = AB_NIL;  // resolved alphabet of the running partition
  bool alphabet_known_ = false;
  int first_guess_ = AB_UNLIMIT;  // plain guess on the first record (stats "type" fallback)
  bool first_block_ = true;
  bool part_fastq_ = false;
  bool any_record_ = false;

  // ---- per-block device state
  const u8 *in_ = nullptr;
  u32 n_ = 0;
  u32 n_nl_ = 0, n_rec_ = 0, n_lines_ = 0;
  bool fastq_ = false;
  bool squeezed_ = false;
  RecIndex ix_{};
  RecArrays ra_{};
  RecViews views_{};
  u32 seq_space_ = 0, qual_space_ = 0;  // size of the byte space seq_off / qual_off index into

  DevStatus *d_status_ = nullptr;
  DevStatus *h_status_ = nullptr;  // pinned
  DevBuf b_in_, b_tile_, b_tile_scan_, b_ls_, b_rl_, b_rec_, b_seq_aoff_, b_qual_aoff_, b_seq_arena_, b_qual_arena_;
  DevBuf b_tmp_, b_keep_, b_out_len_, b_out_off_, b_out_, b_elem_, b_id_, b_gap_seq_, b_gap_qual_, b_newlen_;
  DevBuf b_tables_, b_lens_sorted_, b_rle_u_, b_rle_c_;
  DevBuf b_op1_, b_op2_, b_op3_, b_op4_, b_op5_, b_op6_, b_op7_, b_op8_;
  PinnedBuf h_out_, h_elem_, h_small_;
  static const int kStreamSlots = 3;
  PinnedBuf s_in_slot_[kStreamSlots], s_out_slot_[kStreamSlots];  // pinned block slots of run_stream (kept across calls)
  // bsk_run_buffer pipeline: second input / output buffer sets, copy streams, hand-over events
  DevBuf b_in2_, b_out2_, b_elem2_;
  cudaStream_t s_in_ = nullptr, s_out_ = nullptr;
  cudaEvent_t ev_in_done_[2] = {nullptr, nullptr}, ev_in_free_[2] = {nullptr, nullptr};
  cudaEvent_t ev_out_ready_[2] = {nullptr, nullptr}, ev_out_free_[2] = {nullptr, nullptr};
  u64 launches_ = 0;
  cudaEvent_t ev_[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // start, indexed, main0, main1, end
  bool main_timed_ = false;
  bool contig_known_ = false;  // the caller vouches that every kept record is printed exactly as it stands in the input
  void main_begin();
  void main_end();
  void accumulate_timings();

  // device constant tables (one allocation): class masks, valid[], lut, gap, qual_pow, ...
  u8 *t_class_ = nullptr, *t_valid_ = nullptr, *t_lut_ = nullptr, *t_gap_ = nullptr, *t_aux_ = nullptr;
  double *t_qpow_ = nullptr;

  // ---- stats accumulation (host)
  std::map<u64, u64> hist_;
  u64 q20_ = 0, q30_ = 0, gap_ = 0;
  std::string stats_type_;
  bool stats_type_set_ = false;
  std::vector<u64> hist_len_v_, hist_cnt_v_;

  std::vector<int64_t> keys_host_;
  RmdupState *rm_ = nullptr;
  PatternSet *pats_ = nullptr;
  CommState *comm_ = nullptr;
  DevBuf b_comm_small_, b_comm_, b_fp_, b_fp_before_;
  void comm_gather_words(const u64 *h_mine, size_t n_words, u64 *h_all);
  u64 hit_cap_ = 0;
  int build_patterns(bool only_pos);
  int build_kmer_tables();
  int op_locate_tile(const u8 *d_in, u32 n, int64_t pid, BlockOut &bo);
  int locate_rows(BlockOut &bo, int64_t pid, u64 n_hits);
  int run_matcher(int mode, u8 *flags, u64 &n_hits);
  int finish_grep_count(BlockOut &bo);
  int rmdup_hash_block();
  int rmdup_resolve_block(BlockOut &bo);
  int rmdup_finish(BlockOut &bo, bool prepare_only);
  int rmdup_side_outputs(const u8 *keep, const u64 *first, u64 g_base);
  int op_rmdup_tile(const u8 *d_in, u32 n, BlockOut &bo, bool prepare_only);

  void free_op_state();
  void reset_op_state();
  void reset_status();
  void fetch_status();
  int prepare_block(const u8 *d_in, u32 n);
  int prepare_block_tile(const u8 *d_in, u32 n);
  int resolve_alphabet();
  int check_errors();
  std::string describe_error(u64 rec, u32 kind);
  void upload_tables();
  void set_views_default();

  int process_block(const u8 *d_in, u32 n, int64_t pid, BlockOut &bo);
  int op_seq(BlockOut &bo);
  // single-pass tile kernels for short records (ops_tile.cu); kFusedFallback = the block must take the general path
  static const int kFusedFallback = 1000;
  bool fused_ok_ = true;
  bool inplace_ok_ = true;
  bool seq_fused_eligible() const;
  void seq_emit_cfg(bool fastq, EmitCfg &cfg, u8 *lut, bool &need_lut) const;
  int op_seq_fused(const u8 *d_in, u32 n, BlockOut &bo);
  // same-layout FASTQ kernel (k_fastq_inplace.cu); kFusedFallback when the block is outside its grammar
  int op_seq_inplace(const u8 *d_in, u32 n, bool fastq, const EmitCfg &cfg, const u8 *h_lut, bool need_lut, BlockOut &bo);
  int op_stats_tile(const u8 *d_in, u32 n, BlockOut &bo);
  int n_sm_ = 0;
  u32 part_width_ = 0;  // line width of the running FASTA partition (0 = unwrapped), from its first block
  bool part_width_known_ = false;
  u32 first_rec_bytes_ = 0;
  u32 first_seq_len_ = 0;  // sequence length of the partition's first record (lane-group choice of the tile kernels)
  DevBuf b_tile_cnt_, b_tile_base_, b_slots_;
  int first_record_alphabet(const u8 *d_in, u32 n, bool &fastq, bool &ok, bool long_ok = false);
  PinnedBuf h_probe_;
  int op_stats(BlockOut &bo);
  int op_rmdup(BlockOut &bo, bool prepare_only);
  int op_translate(BlockOut &bo);
  int op_translate_tile(const u8 *d_in, u32 n, BlockOut &bo);
  void translate_host_tables(u8 *tab);
  int op_locate(BlockOut &bo, int64_t pid);
  int op_grep(BlockOut &bo);
  int op_subseq(BlockOut &bo);
  int op_fq2fa(BlockOut &bo);
  int op_duplicate(BlockOut &bo);
  int op_range(BlockOut &bo);
  bool union_ = false;
  u64 range_seen_ = 0;  // records of the partition in front of the running block (RangePrepare index)
  int emit_records(const EmitCfg &cfg, const u8 *keep, const u8 *lut, BlockOut &bo);
  void finalize_stats(bsk_stats *s);
};

int comm_unique_id(uint8_t *id, std::string &err);
int stats_reduce_local(Engine **e, int n, std::string &err);
int rmdup_union_local(Engine **e, int n, const void *const *d_in, const size_t *nbytes, bsk_out *outs, std::string &err);

}  // namespace bsk
