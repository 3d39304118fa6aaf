// ops_basic.cu -- SeqTransform, Stats (+ finalise / render), SubseqTransform and the
// host-buffer streaming entry point.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "engine.h"
#include "prims.h"

namespace bsk {

// ------------------------------------------------------------------ SeqTransform
// bigseqkit-lib/seq.go:81-269 (Call) with the Before() state of :28-79.
int Engine::op_seq(BlockOut &bo) {
  const bool validate = o_.ValidateSeq || !(o_.alphabet == AB_NIL || o_.alphabet == AB_UNLIMIT);  // seq.go:56,66-72
  if (n_rec_ == 0) return BSK_OK;
  if (validate) {
    const u8 *valid = alphabet_valid(alphabet_);
    u8 *h = h_small_.as<u8>();
    memcpy(h, valid, 256);
    BSK_CUDA(cudaMemcpyAsync(t_valid_, h, 256, cudaMemcpyHostToDevice, stream));
    k::validate_seq(views_, t_valid_, (u32)(o_.ValidateSeqLength > 0 ? o_.ValidateSeqLength : 0), d_status_, stream);
    launches_++;
    fetch_status();
  }
  int rc = check_errors();
  if (rc != BSK_OK) return rc;

  // print mode (seq.go:151-163)
  bool print_name = true, print_seq = true, print_qual = fastq_;
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
  const bool qual_only_on_fasta = !fastq_ && !(o_.Name && o_.Seq) && !o_.Name && !o_.Seq && o_.Qual;

  if (o_.RemoveGaps) {  // seq.go:129-131
    u8 *gs = b_gap_seq_.get<u8>((size_t)seq_space_ + 64);
    u8 *gq = b_gap_qual_.get<u8>((size_t)(fastq_ ? qual_space_ : 0) + 64);
    u32 *nl = b_newlen_.get<u32>((size_t)n_rec_ + 1);
    BSK_CUDA(cudaMemsetAsync(nl + n_rec_, 0, 4, stream));
    k::remove_gaps(views_, t_gap_, gs, gq, nl, fastq_ ? 1 : 0, stream);
    launches_++;
    views_.seqb = gs;
    views_.seq_len = nl;
    if (fastq_) {
      views_.qualb = gq;
      views_.qual_len = nl;
    }
  }
  const u8 *keep = nullptr;
  const bool f_len = o_.MinLen > 0 || o_.MaxLen > 0, f_qual = o_.MinQual > 0 || o_.MaxQual > 0;
  if (f_len || f_qual) {  // seq.go:133-149
    u8 *kp = b_keep_.get<u8>((size_t)n_rec_ + 1);
    k::seq_filter(views_, o_.MinLen, o_.MaxLen, o_.MinQual, o_.MaxQual, t_qpow_, kp, stream);
    launches_++;
    keep = kp;
  }
  if (qual_only_on_fasta) {  // seq.go:158-161: raised by the first record that survives the filters
    bool any = n_rec_ > 0;
    if (keep) {
      std::vector<u8> hk(n_rec_);
      BSK_CUDA(cudaMemcpyAsync(hk.data(), keep, n_rec_, cudaMemcpyDeviceToHost, stream));
      BSK_CUDA(cudaStreamSynchronize(stream));
      any = std::find(hk.begin(), hk.end(), (u8)1) != hk.end();
    }
    if (any) {
      err = "FASTA format has no quality. So do not just use flag -q (--qual)";
      return BSK_ERR_DATA;
    }
  }
  if (o_.OnlyId && print_name) {
    u32 *ids = b_id_.get<u32>(((size_t)n_rec_ + 1) * 2);
    k::id_desc(views_, o_.IDNCBI ? 1 : 0, ids, ids + n_rec_ + 1, nullptr, nullptr, stream);
    launches_++;
    views_.name_off = ids;
    views_.name_len = ids + n_rec_ + 1;
  }
  // byte map of the sequence: complement -> dna2rna / rna2dna -> case (seq.go:191-239)
  bool need_lut = false;
  u8 *h = h_small_.as<u8>();
  {
    const u8 *pair = alphabet_pair(alphabet_);
    const bool comp = o_.Complement && alphabet_ != AB_UNLIMIT && alphabet_ != AB_NIL;
    const bool is_rna = alphabet_ == AB_RNA || alphabet_ == AB_RNARED, is_dna = alphabet_ == AB_DNA || alphabet_ == AB_DNARED;
    for (int c = 0; c < 256; c++) {
      u8 x = (u8)c;
      if (comp) x = pair[x];
      if (o_.Dna2rna && !is_rna) x = x == 't' ? 'u' : (x == 'T' ? 'U' : x);
      if (o_.Rna2dna && !is_dna) x = x == 'u' ? 't' : (x == 'U' ? 'T' : x);
      if (o_.LowerCase) { if (x >= 'A' && x <= 'Z') x = (u8)(x + 32); }
      else if (o_.UpperCase) { if (x >= 'a' && x <= 'z') x = (u8)(x - 32); }
      h[c] = x;
      if (x != (u8)c) need_lut = true;
    }
  }
  if (need_lut) {
    BSK_CUDA(cudaMemcpyAsync(t_lut_, h, 256, cudaMemcpyHostToDevice, stream));
    BSK_CUDA(cudaStreamSynchronize(stream));
  }
  EmitCfg cfg;
  cfg.marker = (print_name && print_seq) ? (fastq_ ? '@' : '>') : 0;
  cfg.print_name = print_name;
  cfg.print_seq = print_seq;
  cfg.print_qual = print_qual;
  cfg.plus_line = print_qual && !o_.Qual;
  cfg.reverse = o_.Reverse;
  int width = o_.LineWidth;
  if (o_.Seq || o_.Qual) width = 0;  // seq.go:106-108
  if (fastq_) width = 0;             // seq.go:123
  cfg.width = width > 0 ? (u32)width : 0;
  return emit_records(cfg, keep, need_lut ? t_lut_ : nullptr, bo);
}

// ------------------------------------------------------------------ SubseqTransform, region mode
// bigseqkit-lib/subseq.go:189-190,314-317: Seq.SubSeq(start, end) then Format(LineWidth)
__global__ void k_subseq_region(RecViews v, int start, int end, u32 *seq_off, u32 *seq_len, u32 *qual_off, u32 *qual_len) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= v.n_rec) return;
  const long long len = v.seq_len[r];
  long long s = start, e = end;
  u32 s0 = 0, sl = 0;
  bool ok = len > 0;
  if (ok) {  // seq.SubLocation of bio v0.7.0, pinned by the table in bigseqkit-cli/helper.go:348-361
    if (s < 1) {
      if (s == 0) s = 1;
      else if (e < 0 && s > e) ok = false;
      else s = (-s > len) ? 1 : len + s + 1;
    } else if (s > len) ok = false;
  }
  if (ok) {
    if (e > len) e = len;
    else if (e < 1) {
      if (e == 0) e = -1;
      if (-e > len) ok = false;
      else e = len + e + 1;
    }
  }
  if (ok && s - 1 > e) ok = false;
  if (ok) {
    s0 = (u32)(s - 1);
    sl = (u32)(e - (s - 1));
  }
  seq_off[r] = v.seq_off[r] + s0;
  seq_len[r] = sl;
  const u32 ql = v.qual_len[r];
  qual_off[r] = v.qual_off[r] + (ql ? s0 : 0);
  qual_len[r] = ql ? sl : 0;
}

int Engine::op_subseq(BlockOut &bo) {
  if (n_rec_ == 0) return BSK_OK;
  int rc = check_errors();
  if (rc != BSK_OK) return rc;
  const size_t R = (size_t)n_rec_ + 1;
  u32 *a = b_op1_.get<u32>(R * 4);
  BSK_LAUNCH_FLAT(k_subseq_region, (n_rec_ + 255) / 256, 256, 0, stream, views_, o_.region_start, o_.region_end, a, a + R,
                  a + 2 * R, a + 3 * R);
  launches_++;
  views_.seq_off = a;
  views_.seq_len = a + R;
  views_.qual_off = a + 2 * R;
  views_.qual_len = a + 3 * R;
  EmitCfg cfg;
  cfg.marker = fastq_ ? '@' : '>';
  cfg.print_name = 1;
  cfg.print_seq = 1;
  cfg.print_qual = fastq_;
  cfg.plus_line = fastq_;
  cfg.reverse = 0;
  cfg.width = fastq_ ? 0 : (o_.LineWidth > 0 ? (u32)o_.LineWidth : 0);
  return emit_records(cfg, nullptr, nullptr, bo);
}

// ------------------------------------------------------------------ Fq2Fa
// Fq2Fa.Call (bigseqkit-lib/fq2fa.go:36-61): qualities dropped, Record.Format(0) minus the final '\n' -> ">Name\nseq"
// on one line whatever the input's wrapping; FASTA input goes through the same formatter.
int Engine::op_fq2fa(BlockOut &bo) {
  int rc = check_errors();
  if (rc != BSK_OK) return rc;
  EmitCfg cfg{};
  cfg.marker = '>';
  cfg.print_name = 1;
  cfg.print_seq = 1;
  return emit_records(cfg, nullptr, nullptr, bo);
}

// ------------------------------------------------------------------ Duplicate, Range (Head)
// Operators on the RAW elements: record text as it stands minus one trailing '\n' (ReadFixer, bigseqkit-lib/helper.go:51),
// written back with FileStore's '\n' -- i.e. the bytes of the record, plus a '\n' behind a last record that has none.
//   Duplicate.Call      bigseqkit-lib/duplicate.go:24-30   every element `times` times
//   RangePrepare.Call   bigseqkit-lib/range.go:26-31       kept iff start <= index < end (index: MapWithIndex, global)
//   RangeFilter.Call    bigseqkit-lib/range.go:42-44 ; Head = Range "1:N" (bigseqkit/head.go:33-44)
__global__ void k_dup_copy(const u8 *__restrict__ in, u32 n, const u32 *__restrict__ head_off, u32 n_rec, u32 times, int add_nl,
                           u8 *__restrict__ out, u64 *__restrict__ elem_off) {
  const u32 r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
  if (r >= n_rec) return;
  const u32 s = head_off[r] - 1u, e = r + 1 < n_rec ? head_off[r + 1] - 1u : n;
  const u32 len = e - s;
  const u64 elen = (u64)len + ((r + 1 == n_rec && add_nl) ? 1u : 0u);
  const u64 base = (u64)times * s;  // every record in front of this one ends with its own '\n'
  for (u32 c = 0; c < times; c++) {
    u8 *dst = out + base + c * elen;
    for (u32 i = lane; i < len; i += 32) dst[i] = in[s + i];
    if (lane == 0) {
      if (elen > len) dst[len] = '\n';
      if (elem_off) elem_off[(u64)r * times + c] = base + c * elen;
    }
  }
}
__global__ void k_range_elems(const u32 *__restrict__ head_off, u32 a, u32 b, u32 s_a, u64 total, u64 *__restrict__ elem_off) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > b - a) return;
  elem_off[i] = i == b - a ? total : (u64)(head_off[a + i] - 1u - s_a);
}

int Engine::op_duplicate(BlockOut &bo) {
  // (no check_errors(): the reference does not parse the records here, whatever they hold is passed on)
  if (!n_rec_ || o_.Times == 0) return BSK_OK;
  if (o_.Times > 0xffffffffll) { err = "duplicate: times is too large"; return BSK_ERR_ARG; }
  u8 *hs = h_small_.as<u8>();
  prim::copy_small(hs, in_ + n_ - 1, 1, stream);
  BSK_CUDA(cudaStreamSynchronize(stream));
  const int add_nl = hs[0] != '\n';
  const u64 total = (u64)o_.Times * ((u64)n_ + (add_nl ? 1u : 0u));
  const u64 n_el = (u64)n_rec_ * (u64)o_.Times;
  u8 *out = b_out_.get<u8>((size_t)total + 64);
  u64 *elem = want_elem_off ? b_elem_.get<u64>((size_t)n_el + 1) : nullptr;
  main_begin();
  BSK_LAUNCH_FLAT(k_dup_copy, (u32)(((u64)n_rec_ * 32 + 255) / 256), 256, 0, stream, in_, n_, ra_.head_off, n_rec_, (u32)o_.Times,
                  add_nl, out, elem);
  main_end();
  launches_++;
  if (elem) {
    memcpy(hs, &total, 8);
    BSK_CUDA(cudaMemcpyAsync(elem + n_el, hs, 8, cudaMemcpyHostToDevice, stream));
  }
  bo.d_data = out;
  bo.n = total;
  bo.d_elem_off = elem;
  bo.n_elem = n_el;
  return BSK_OK;
}

int Engine::op_range(BlockOut &bo) {
  if (first_block_ && !union_) range_seen_ = 0;
  const int64_t idx0 = o_.IndexBase + (int64_t)range_seen_;  // global index of this block's first record
  range_seen_ += n_rec_;
  int64_t a = o_.RangeStart - idx0, b = o_.RangeEnd > idx0 + (int64_t)n_rec_ ? (int64_t)n_rec_ : o_.RangeEnd - idx0;
  if (a < 0) a = 0;
  if (b > (int64_t)n_rec_) b = n_rec_;
  if (b <= a) return BSK_OK;
  u32 *hs = h_small_.as<u32>();
  prim::copy_small(hs, ra_.head_off + a, 4, stream);
  if (b < (int64_t)n_rec_) prim::copy_small(hs + 1, ra_.head_off + b, 4, stream);
  prim::copy_small(hs + 2, in_ + n_ - 1, 1, stream);
  BSK_CUDA(cudaStreamSynchronize(stream));
  const u32 s_a = hs[0] - 1u, s_b = b < (int64_t)n_rec_ ? hs[1] - 1u : n_;
  const bool add_nl = b == (int64_t)n_rec_ && *reinterpret_cast<const u8 *>(hs + 2) != '\n';
  const u64 total = (u64)(s_b - s_a) + (add_nl ? 1u : 0u);
  u8 *out = b_out_.get<u8>((size_t)total + 64);
  main_begin();
  BSK_CUDA(cudaMemcpyAsync(out, in_ + s_a, s_b - s_a, cudaMemcpyDeviceToDevice, stream));
  if (add_nl) BSK_CUDA(cudaMemsetAsync(out + total - 1, '\n', 1, stream));
  main_end();
  const u32 n_el = (u32)(b - a);
  u64 *elem = nullptr;
  if (want_elem_off) {
    elem = b_elem_.get<u64>((size_t)n_el + 1);
    BSK_LAUNCH_FLAT(k_range_elems, (n_el + 1 + 255) / 256, 256, 0, stream, ra_.head_off, (u32)a, (u32)b, s_a, total, elem);
    launches_++;
  }
  bo.d_data = out;
  bo.n = total;
  bo.d_elem_off = elem;
  bo.n_elem = n_el;
  return BSK_OK;
}

// ------------------------------------------------------------------ Stats
// bigseqkit-lib/stats.go:48-117 (per partition) ; totals are kept with sum semantics
int Engine::op_stats(BlockOut &bo) {
  (void)bo;
  int rc = check_errors();
  if (rc != BSK_OK) return rc;
  if (n_rec_ == 0) return BSK_OK;
  u32 *sorted = b_lens_sorted_.get<u32>(n_rec_);
  u32 *uq = b_rle_u_.get<u32>(n_rec_);
  u32 *cn = b_rle_c_.get<u32>(n_rec_);
  prim::sort_u32(views_.seq_len, sorted, n_rec_, b_tmp_, stream);
  prim::rle_u32(sorted, uq, cn, &d_status_->n_sel, n_rec_, b_tmp_, stream);
  if (o_.All) {
    k::stats_qual_gap(views_, t_gap_, o_.fq_offset, fastq_ ? 1 : 0, d_status_, stream);
    launches_++;
  }
  fetch_status();
  const u32 runs = h_status_->n_sel;
  std::vector<u32> hu(runs), hc(runs);
  if (runs) {
    BSK_CUDA(cudaMemcpyAsync(hu.data(), uq, runs * 4ull, cudaMemcpyDeviceToHost, stream));
    BSK_CUDA(cudaMemcpyAsync(hc.data(), cn, runs * 4ull, cudaMemcpyDeviceToHost, stream));
    BSK_CUDA(cudaStreamSynchronize(stream));
  }
  for (u32 i = 0; i < runs; i++) hist_[hu[i]] += hc[i];
  if (o_.All) {
    q20_ += h_status_->counters[0];
    q30_ += h_status_->counters[1];
    gap_ += h_status_->counters[2];
  }
  if (!stats_type_set_) {  // stats.go:106-114 + bigseqkit/stats.go:109-130
    if (alphabet_ == AB_DNARED) stats_type_ = "DNA";
    else if (alphabet_ == AB_RNARED) stats_type_ = "RNA";
    else stats_type_ = alphabet_name(first_guess_);
    stats_type_set_ = true;
  }
  return BSK_OK;
}

static double round_n(double f, int n) {  // util/math.Round of shenwei356/util v0.5.0
  const double p = pow(10, n);
  return trunc((f + 0.5 / p) * p) / p;
}

namespace {
struct HistView {
  const std::vector<u64> &len, &cnt;
  u64 at(u64 idx) const {  // idx-th smallest length
    u64 c = 0;
    for (size_t i = 0; i < len.size(); i++) {
      c += cnt[i];
      if (idx < c) return len[i];
    }
    return 0;
  }
  double mid(u64 a, u64 b) const { return ((double)at(a) + (double)at(b)) / 2; }
};
}  // namespace

// bio/util.LengthStats + bigseqkit/stats.go:132-161
void Engine::finalize_stats(bsk_stats *s) {
  memset(s, 0, sizeof *s);
  hist_len_v_.clear();
  hist_cnt_v_.clear();
  for (auto &kv : hist_) {
    hist_len_v_.push_back(kv.first);
    hist_cnt_v_.push_back(kv.second);
  }
  s->hist_len = hist_len_v_.data();
  s->hist_cnt = hist_cnt_v_.data();
  s->n_hist = hist_len_v_.size();
  std::string type = stats_type_;
  if (!stats_type_set_) {
    const int a = o_.alphabet == AB_NIL ? AB_UNLIMIT : o_.alphabet;
    type = a == AB_DNARED ? "DNA" : a == AB_RNARED ? "RNA" : a == AB_UNLIMIT ? "" : alphabet_name(AB_UNLIMIT);
  }
  snprintf(s->type, sizeof s->type, "%s", type.c_str());
  s->q20 = q20_;
  s->q30 = q30_;
  u64 num = 0, sum = 0;
  for (size_t i = 0; i < s->n_hist; i++) {
    num += hist_cnt_v_[i];
    sum += hist_len_v_[i] * hist_cnt_v_[i];
  }
  s->num = num;
  s->sum_len = sum;
  if (num == 0) return;  // all zeros (stats.go:149-161)
  s->sum_gap = gap_;
  s->min_len = hist_len_v_.front();
  s->max_len = hist_len_v_.back();
  s->avg_len = round_n((double)sum / (double)num, 1);
  if (o_.All) {
    HistView hv{hist_len_v_, hist_cnt_v_};
    const double half = (double)sum / 2;
    double acc = 0;
    u64 l50 = 0;
    for (size_t i = s->n_hist; i-- > 0;) {
      acc += (double)(hist_len_v_[i] * hist_cnt_v_[i]);
      l50 += hist_cnt_v_[i];
      if (acc >= half) {
        s->n50 = hist_len_v_[i];
        s->l50 = l50;
        break;
      }
    }
    const bool even = (num & 1) == 0;
    s->q2 = even ? hv.mid(num / 2 - 1, num / 2) : (double)hv.at(num / 2);
    const u64 h = even ? num / 2 : (num + 1) / 2, m = num / 2;
    const bool heven = (h % 2) == 0;
    s->q1 = heven ? hv.mid(h / 2 - 1, h / 2) : (double)hv.at(h / 2);
    s->q3 = heven ? hv.mid(m + h / 2 - 1, m + h / 2) : (double)hv.at(m + h / 2);
  }
  s->q20_pct = round_n((double)q20_ / (double)sum * 100, 2);
  s->q30_pct = round_n((double)q30_ / (double)sum * 100, 2);
}

int Engine::stats_result(bsk_stats *out) {
  if (op_ != OP_STATS) { err = "bsk_stats_result: ctx is not a Stats operator"; return BSK_ERR_STATE; }
  finalize_stats(out);
  return BSK_OK;
}

int Engine::stats_add(const u64 *len, const u64 *cnt, size_t n, u64 q20, u64 q30, u64 gap, const char *type) {
  if (op_ != OP_STATS) { err = "bsk_stats_add: ctx is not a Stats operator"; return BSK_ERR_STATE; }
  for (size_t i = 0; i < n; i++) hist_[len[i]] += cnt[i];
  q20_ += q20;
  q30_ += q30;
  gap_ += gap;
  if (!stats_type_set_ && type && (n > 0 || type[0])) {
    stats_type_ = type;
    stats_type_set_ = true;
  }
  return BSK_OK;
}

int Engine::stats_merge_from(const Engine &src) {  // StatsReduce (bigseqkit-lib/stats.go:128-137), sum semantics
  if (op_ != OP_STATS || src.op_ != OP_STATS) { err = "bsk_stats_merge: both ctx must be Stats operators"; return BSK_ERR_STATE; }
  for (auto &kv : src.hist_) hist_[kv.first] += kv.second;
  q20_ += src.q20_;
  q30_ += src.q30_;
  gap_ += src.gap_;
  if (!stats_type_set_ && src.stats_type_set_) {
    stats_type_ = src.stats_type_;
    stats_type_set_ = true;
  }
  return BSK_OK;
}

__global__ void k_dense_hist_fill(const u64 *len, const u64 *cnt, u32 n, u64 *hist, u64 nbins) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && len[i] < nbins) hist[len[i]] = cnt[i];
}

int Engine::stats_dense_device(void *d_hist, size_t nbins, u64 *n_overflow) {
  if (op_ != OP_STATS) { err = "bsk_stats_dense_device: ctx is not a Stats operator"; return BSK_ERR_STATE; }
  if (device_ >= 0) BSK_CUDA(cudaSetDevice(device_));
  std::vector<u64> l, c;
  u64 over = 0;
  for (auto &kv : hist_) {
    if (kv.first < nbins) { l.push_back(kv.first); c.push_back(kv.second); }
    else over += kv.second;
  }
  BSK_CUDA(cudaMemsetAsync(d_hist, 0, nbins * sizeof(u64), stream));
  if (!l.empty()) {
    u64 *dl = b_op1_.get<u64>(l.size() * 2);
    BSK_CUDA(cudaMemcpyAsync(dl, l.data(), l.size() * 8, cudaMemcpyHostToDevice, stream));
    BSK_CUDA(cudaMemcpyAsync(dl + l.size(), c.data(), l.size() * 8, cudaMemcpyHostToDevice, stream));
    BSK_LAUNCH_FLAT(k_dense_hist_fill, (u32)((l.size() + 255) / 256), 256, 0, stream, dl, dl + l.size(), (u32)l.size(),
                    static_cast<u64 *>(d_hist), (u64)nbins);
  }
  BSK_CUDA(cudaStreamSynchronize(stream));
  if (n_overflow) *n_overflow = over;
  return BSK_OK;
}

// humanize.Comma / Commaf (dustin/go-humanize, used by bigseqkit/stats.go:227-285)
static std::string comma_u(u64 v) {
  char t[32];
  const int n = snprintf(t, sizeof t, "%llu", (unsigned long long)v);
  std::string o;
  for (int i = 0; i < n; i++) {
    o += t[i];
    if ((n - 1 - i) % 3 == 0 && i != n - 1) o += ',';
  }
  return o;
}
static std::string comma_f(double v) {
  char t[64];
  snprintf(t, sizeof t, "%.10g", v);
  const char *dot = strchr(t, '.');
  std::string ip(t, dot ? (size_t)(dot - t) : strlen(t));
  std::string o = comma_u(strtoull(ip.c_str(), nullptr, 10));
  if (dot) o += dot;
  return o;
}

// StatsString (bigseqkit/stats.go:168-288)
long Engine::stats_render(const char *file, const char *format, char *buf, size_t cap) {
  bsk_stats s;
  finalize_stats(&s);
  std::string b;
  char t[512];
  const bool all = o_.All;
  if (o_.Tabular) {
    b += "file\tformat\ttype\tnum_seqs\tsum_len\tmin_len\tavg_len\tmax_len";
    if (all) b += "\tQ1\tQ2\tQ3\tsum_gap\tN50\tQ20(%)\tQ30(%)";
    b += '\n';
    snprintf(t, sizeof t, "%s\t%s\t%s\t%llu\t%llu\t%llu\t%.1f\t%llu", file, format, s.type, (unsigned long long)s.num,
             (unsigned long long)s.sum_len, (unsigned long long)s.min_len, s.avg_len, (unsigned long long)s.max_len);
    b += t;
    if (all) {
      snprintf(t, sizeof t, "\t%.1f\t%.1f\t%.1f\t%llu\t%llu\t%.2f\t%.2f", s.q1, s.q2, s.q3, (unsigned long long)s.sum_gap,
               (unsigned long long)s.n50, s.q20_pct, s.q30_pct);
      b += t;
    }
    b += '\n';
  } else {
    static const char *hdr[15] = {"file", "format", "type", "num_seqs", "sum_len", "min_len", "avg_len", "max_len",
                                  "Q1", "Q2", "Q3", "sum_gap", "N50", "Q20(%)", "Q30(%)"};
    std::string cell[15];
    const int nc = all ? 15 : 8;
    cell[0] = file; cell[1] = format; cell[2] = s.type;
    cell[3] = comma_u(s.num); cell[4] = comma_u(s.sum_len); cell[5] = comma_u(s.min_len);
    cell[6] = comma_f(s.avg_len); cell[7] = comma_u(s.max_len);
    if (all) {
      cell[8] = comma_f(s.q1); cell[9] = comma_f(s.q2); cell[10] = comma_f(s.q3);
      cell[11] = comma_u(s.sum_gap); cell[12] = comma_u(s.n50); cell[13] = comma_f(s.q20_pct); cell[14] = comma_f(s.q30_pct);
    }
    for (int row = 0; row < 2; row++) {
      for (int c = 0; c < nc; c++) {
        const std::string txt = row == 0 ? hdr[c] : cell[c];
        const size_t w = std::max(strlen(hdr[c]), cell[c].size());
        const size_t pad = w - txt.size();
        const bool right = c >= 3;
        if (c) b += ' ';
        if (right) b.append(pad, ' ');
        b += txt;
        if (!right) b.append(pad, ' ');
      }
      b += '\n';
    }
  }
  if (buf && cap) {
    const size_t k = std::min(cap - 1, b.size());
    memcpy(buf, b.data(), k);
    buf[k] = 0;
  }
  return (long)b.size();
}

// ------------------------------------------------------------------ host-buffer entry point
// One Call() on a partition in host memory: cut it into record-aligned blocks of at most BSK_BLOCK_BYTES
// (default 64 MiB), and run them through a three-stream pipeline so that PCIe stays busy in both directions
// while the kernels run:
//   copy-in stream   H2D of block i+1 into the other input buffer
//   ctx stream       kernels of block i
//   copy-out stream  D2H of block i-1 (records + element offsets) from the other output buffer
// With pinned host memory the copies are true DMA; the end-to-end rate is then bound by PCIe, not by the kernels.
static size_t env_block_bytes() {
  const char *e = getenv("BSK_BLOCK_BYTES");
  size_t v = e ? strtoull(e, nullptr, 10) : 0;
  if (v < 4096) v = 64ull << 20;
  if (v > kMaxBlockBytes / 2) v = kMaxBlockBytes / 2;
  return v;
}

static size_t next_record_start(const u8 *d, size_t n, size_t from, bool fq) {
  const u8 marker = fq ? '@' : '>';
  for (size_t L = from < 1 ? 1 : from; L < n; L++) {
    if (d[L - 1] != '\n' || d[L] != marker) continue;
    if (fq && L >= 3 && d[L - 3] == '\n' && d[L - 2] == '+') continue;
    return L;
  }
  return n;
}

__global__ void k_add_u64(u64 *p, u64 n, u64 add) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] += add;
}

int Engine::run_buffer(const u8 *in, size_t n, int64_t pid, bsk_out *out) {
  memset(out, 0, sizeof *out);
  if (device_ >= 0) BSK_CUDA(cudaSetDevice(device_));
  launches_ = 0;
  timings = bsk_timings{};
  alphabet_ = o_.alphabet;
  alphabet_known_ = false;
  first_block_ = true;
  if (!s_in_) {
    BSK_CUDA(cudaStreamCreateWithFlags(&s_in_, cudaStreamNonBlocking));
    BSK_CUDA(cudaStreamCreateWithFlags(&s_out_, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      BSK_CUDA(cudaEventCreateWithFlags(&ev_in_done_[i], cudaEventDisableTiming));
      BSK_CUDA(cudaEventCreateWithFlags(&ev_in_free_[i], cudaEventDisableTiming));
      BSK_CUDA(cudaEventCreateWithFlags(&ev_out_ready_[i], cudaEventDisableTiming));
      BSK_CUDA(cudaEventCreateWithFlags(&ev_out_free_[i], cudaEventDisableTiming));
    }
  }
  // ---- record-aligned cut points
  const size_t blk = env_block_bytes();
  const bool fq = n > 0 && in[0] == '@';
  std::vector<size_t> cut;
  cut.push_back(0);
  while (cut.back() < n) {
    const size_t pos = cut.back();
    size_t end = n;
    if (n - pos > blk) {
      end = next_record_start(in, n, pos + blk, fq);
      if (end - pos >= kMaxBlockBytes) { err = "a single record exceeds the 4 GiB block limit"; return BSK_ERR_DATA; }
    }
    cut.push_back(end);
  }
  if (cut.size() == 1) cut.push_back(0);  // empty partition: one empty block
  const size_t nb = cut.size() - 1;
  size_t max_block = 0;
  for (size_t i = 0; i < nb; i++) max_block = std::max(max_block, cut[i + 1] - cut[i]);
  u8 *d_in[2] = {b_in_.get<u8>(max_block + 64), nb > 1 ? b_in2_.get<u8>(max_block + 64) : nullptr};
  // outputs rarely exceed the input by much; growing later costs a pipeline drain
  h_out_.reserve(n + n / 16 + 4096, false, 0);
  if (want_elem_off) h_elem_.reserve(4096, false, 0);

  auto upload = [&](size_t i) {
    const size_t bn = cut[i + 1] - cut[i];
    if (i >= 2) BSK_CUDA(cudaStreamWaitEvent(s_in_, ev_in_free_[i & 1], 0));  // kernels of block i-2 are done with it
    if (bn) BSK_CUDA(cudaMemcpyAsync(d_in[i & 1], in + cut[i], bn, cudaMemcpyHostToDevice, s_in_));
    BSK_CUDA(cudaEventRecord(ev_in_done_[i & 1], s_in_));
  };

  size_t out_used = 0, elem_used = 0;
  u64 n_rec_total = 0;
  bool out_busy[2] = {false, false};  // an asynchronous D2H still reads output buffer set b
  int ob = 0;                         // output buffer set the next block writes
  upload(0);
  for (size_t i = 0; i < nb; i++) {
    if (i + 1 < nb) upload(i + 1);
    const size_t bn = cut[i + 1] - cut[i];
    BSK_CUDA(cudaStreamWaitEvent(stream, ev_in_done_[i & 1], 0));
    if (out_busy[ob]) BSK_CUDA(cudaStreamWaitEvent(stream, ev_out_free_[ob], 0));
    BlockOut bo;
    int rc = process_block(d_in[i & 1], (u32)bn, pid, bo);  // synchronises the ctx stream
    if (rc != BSK_OK) {
      cudaStreamSynchronize(s_in_);
      cudaStreamSynchronize(s_out_);
      return rc;
    }
    BSK_CUDA(cudaEventRecord(ev_in_free_[i & 1], stream));
    n_rec_total += bo.n_rec;
    // host arenas: growing them moves the data, so drain the copy-out stream first
    if (out_used + bo.n + 64 > h_out_.cap) {
      BSK_CUDA(cudaStreamSynchronize(s_out_));
      h_out_.reserve(out_used + bo.n + 64, true, out_used);
    }
    if (want_elem_off && (elem_used + bo.n_elem + 2) * 8 > h_elem_.cap) {
      BSK_CUDA(cudaStreamSynchronize(s_out_));
      h_elem_.reserve((elem_used + bo.n_elem + 2) * 8, true, elem_used * 8);
    }
    const bool want_elems = want_elem_off && bo.n_elem && bo.d_elem_off;
    if (want_elems && out_used) {  // element offsets are relative to the block's output
      BSK_LAUNCH_FLAT(k_add_u64, (u32)((bo.n_elem + 255) / 256), 256, 0, stream, bo.d_elem_off, (u64)bo.n_elem, (u64)out_used);
      launches_++;
    }
    BSK_CUDA(cudaEventRecord(ev_out_ready_[ob], stream));
    BSK_CUDA(cudaStreamWaitEvent(s_out_, ev_out_ready_[ob], 0));
    if (bo.n) BSK_CUDA(cudaMemcpyAsync(h_out_.as<u8>() + out_used, bo.d_data, bo.n, cudaMemcpyDeviceToHost, s_out_));
    if (want_elems)
      BSK_CUDA(cudaMemcpyAsync(h_elem_.as<u64>() + elem_used, bo.d_elem_off, bo.n_elem * 8, cudaMemcpyDeviceToHost, s_out_));
    BSK_CUDA(cudaEventRecord(ev_out_free_[ob], s_out_));
    // the next block may only run ahead of this D2H if its results land in other buffers
    const bool swappable = (bo.n == 0 || bo.d_data == b_out_.p) && (!want_elems || bo.d_elem_off == b_elem_.p);
    if (swappable && i + 1 < nb) {
      std::swap(b_out_.p, b_out2_.p);
      std::swap(b_out_.cap, b_out2_.cap);
      std::swap(b_elem_.p, b_elem2_.p);
      std::swap(b_elem_.cap, b_elem2_.cap);
      out_busy[ob] = true;
      ob ^= 1;
    } else {
      BSK_CUDA(cudaStreamSynchronize(s_out_));
      out_busy[ob] = false;
    }
    out_used += bo.n;
    elem_used += bo.n_elem;
  }
  BSK_CUDA(cudaStreamSynchronize(s_out_));
  BSK_CUDA(cudaStreamSynchronize(s_in_));
  if (op_ == OP_GREP && o_.Count) {
    BlockOut bo;
    int rc = finish_grep_count(bo);
    if (rc != BSK_OK) return rc;
    h_out_.reserve(bo.n + 64, false, 0);
    BSK_CUDA(cudaMemcpy(h_out_.as<u8>(), bo.d_data, bo.n, cudaMemcpyDeviceToHost));
    out_used = bo.n;
    elem_used = 1;
    if (want_elem_off) {
      h_elem_.reserve(4 * 8, false, 0);
      h_elem_.as<u64>()[0] = 0;
    }
  }
  if (want_elem_off) {
    h_elem_.reserve((elem_used + 2) * 8, true, elem_used * 8);
    h_elem_.as<u64>()[elem_used] = out_used;
    out->elem_off = h_elem_.as<u64>();
  }
  h_out_.reserve(out_used + 64, true, out_used);
  out->data = h_out_.as<u8>();
  out->n = out_used;
  out->n_elem = elem_used;
  out->n_records = n_rec_total;
  timings.kernel_launches = launches_;
  timings.in_bytes = n;
  timings.out_bytes = out_used;
  return BSK_OK;
}

}  // namespace bsk
