// ops_rmdup.cu -- rmdup: subject hashing, first-occurrence resolution, survivor emission.
//
//   RmDupPrepare.Call   bigseqkit-lib/rmdup.go:43-90    key = int64(xxhash.Sum64(subject))
//   GroupByKey          bigseqkit/rmdup.go:97           -> one device hash table  key -> earliest record
//   RmDupCheck.Call     bigseqkit-lib/rmdup.go:118-242  exact-subject compare inside an equal-key group, first wins
// Pinned semantics: the first occurrence in input order survives, output in input order.
#include <algorithm>
#include <cstring>

#include "engine.h"
#include "op_state.h"
#include "prims.h"
#include "xxh64.cuh"

namespace bsk {

static const u64 kNoFirst = ~0ull;

struct SubjectViews {
  const u8 *base;
  const u32 *off, *len;
  u64 limit;  // readable bytes of base (word loads stop there)
};

struct GetRaw {
  const u8 *p;
  __host__ __device__ __forceinline__ u8 operator()(u32 i) const { return p[i]; }
};
struct GetLower {
  const u8 *p;
  __host__ __device__ __forceinline__ u8 operator()(u32 i) const {
    const u8 c = p[i];
    return (c >= 'A' && c <= 'Z') ? (u8)(c + 32) : c;  // bytes.ToLower on ASCII
  }
};

// one thread per record
__global__ void k_rmdup_hash(SubjectViews sv, u32 n_rec, int ignore_case, u64 *__restrict__ keys, u64 *__restrict__ fps) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rec) return;
  const u8 *p = sv.base + sv.off[r];
  const u32 len = sv.len[r];
  u64 a, b;
  if (ignore_case) xxh64_key_fp(GetLower{p}, len, a, b);
  else xxh64_key_fp(GetRaw{p}, len, a, b);
  keys[r] = a;
  fps[r] = b;
}

// open addressing, linear probing over 16-byte slots {key, earliest ordinal}: a probe touches one sector.  An empty
// slot is all ones (one memset clears the table and sets every ordinal to "none"); that key lives in the extra slot `cap`.
static const u64 kEmptyKey = ~0ull;

__device__ __forceinline__ u64 table_insert_slot(TableSlot *t, u64 cap, u64 key) {
  if (key == kEmptyKey) return cap;
  u64 i = (key * 0x9E3779B97F4A7C15ull) >> 20 & (cap - 1);
  for (;;) {
    const u64 cur = t[i].key;
    if (cur == key) return i;
    if (cur == kEmptyKey) {
      const u64 old = atomicCAS((unsigned long long *)&t[i].key, (unsigned long long)kEmptyKey, (unsigned long long)key);
      if (old == kEmptyKey || old == key) return i;
    }
    i = (i + 1) & (cap - 1);
  }
}

// earliest ordinal stored for `key` (the key is known to be in the table)
__device__ __forceinline__ u64 table_first(const TableSlot *t, u64 cap, u64 key) {
  if (key == kEmptyKey) return t[cap].first;
  u64 i = (key * 0x9E3779B97F4A7C15ull) >> 20 & (cap - 1);
  for (;;) {
    const ulonglong2 e = *reinterpret_cast<const ulonglong2 *>(&t[i]);
    if (e.x == key) return e.y;
    if (e.x == kEmptyKey) return kNoFirst;
    i = (i + 1) & (cap - 1);
  }
}

__global__ void k_table_insert(const u64 *__restrict__ keys, u64 n, u64 g_base, TableSlot *t, u64 cap) {
  const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const u64 s = table_insert_slot(t, cap, keys[r]);
  atomicMin((unsigned long long *)&t[s].first, (unsigned long long)(g_base + r));
}

__device__ __forceinline__ bool subject_equal(const SubjectViews &sv, u32 a, u32 b, int ignore_case) {
  const u32 la = sv.len[a];
  if (la != sv.len[b]) return false;
  const u32 oa = sv.off[a], ob = sv.off[b];
  const u8 *pa = sv.base + oa, *pb = sv.base + ob;
  if (ignore_case) {
    GetLower ga{pa}, gb{pb};
    for (u32 i = 0; i < la; i++)
      if (ga(i) != gb(i)) return false;
    return true;
  }
  const u64 top = (u64)(oa > ob ? oa : ob) + la + 8u;
  if (top > sv.limit || (((size_t)sv.base) & 3u)) {  // too close to the end of the buffer for word loads
    for (u32 i = 0; i < la; i++)
      if (pa[i] != pb[i]) return false;
    return true;
  }
  // four bytes per step: aligned 32-bit loads, funnel shifts resolve the two alignments
  const u32 *wa = reinterpret_cast<const u32 *>(sv.base + (oa & ~3u)), *wb = reinterpret_cast<const u32 *>(sv.base + (ob & ~3u));
  const u32 sa = (oa & 3u) * 8u, sb = (ob & 3u) * 8u;
  u32 a0 = wa[0], b0 = wb[0];
  const u32 nw = (la + 3u) >> 2;
  for (u32 i = 0; i < nw; i++) {
    const u32 a1 = wa[i + 1], b1 = wb[i + 1];
    u32 x = __funnelshift_r(a0, a1, sa) ^ __funnelshift_r(b0, b1, sb);
    if (i + 1 == nw && (la & 3u)) x &= (1u << (8u * (la & 3u))) - 1u;
    if (x) return false;
    a0 = a1;
    b0 = b1;
  }
  return true;
}

// keep[r]: 1 first occurrence, 0 duplicate, 2 unresolved (64-bit key collision between different subjects)
__global__ void k_rmdup_resolve(SubjectViews sv, u32 n_rec, int ignore_case, const u64 *__restrict__ keys,
                                const u64 *__restrict__ fps, u64 g_base, const u64 *__restrict__ hist_fp,
                                const TableSlot *__restrict__ table, u64 cap, u8 *__restrict__ keep,
                                u64 *__restrict__ first_out, DevStatus *st) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rec) return;
  const u64 g = g_base + r;
  const u64 first = table_first(table, cap, keys[r]);
  u8 k;
  if (first == g) k = 1;
  else if (first >= g_base) k = subject_equal(sv, r, (u32)(first - g_base), ignore_case) ? 0 : 2;
  else k = (hist_fp[first] == fps[r]) ? 0 : 2;  // earlier block / other GPU: second 64-bit hash decides
  keep[r] = k;
  if (first_out) first_out[r] = first;  // ordinal of the group's first member (rmdup -D)
  if (k == 2) atomicAdd((unsigned long long *)&st->counters[4], 1ull);
}

// rare path: serial scan for the records whose key collided with a different subject
__global__ void k_rmdup_fixup(SubjectViews sv, u32 n_rec, int ignore_case, const u64 *keys, const u64 *fps, u64 g_base,
                              const u64 *hist_keys, const u64 *hist_fp, u8 *keep, u64 *first_out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (u32 r = 0; r < n_rec; r++) {
    if (keep[r] != 2) continue;
    u64 first = g_base + r;
    bool dup = false;
    for (u64 h = 0; h < g_base && !dup; h++)
      if (hist_keys[h] == keys[r] && hist_fp[h] == fps[r]) { dup = true; first = h; }
    for (u32 q = 0; q < r && !dup; q++)
      if (keys[q] == keys[r] && subject_equal(sv, r, q, ignore_case)) { dup = true; first = g_base + q; }
    keep[r] = dup ? 0 : 1;
    if (first_out) first_out[r] = first;
  }
}

__global__ void k_rmdup_drop_mask(const u8 *__restrict__ keep, u32 n_rec, u8 *__restrict__ drop) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rec) drop[r] = keep[r] == 0;
}

__global__ void k_interleave_fp(const u64 *keys, const u64 *fps, u64 n, u64 *out) {
  const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  out[2 * r] = keys[r];
  out[2 * r + 1] = fps[r];
}
__global__ void k_deinterleave_fp(const u64 *in, u64 n, u64 *keys, u64 *fps) {
  const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  keys[r] = in[2 * r];
  fps[r] = in[2 * r + 1];
}

// ------------------------------------------------------------------ host side
static void hist_reserve(Engine::RmdupState *rm, u64 need, cudaStream_t s) {
  if (need <= rm->hist_cap) return;
  u64 cap = rm->hist_cap ? rm->hist_cap : (1u << 16);
  while (cap < need) cap *= 2;
  u64 *nk = nullptr, *nf = nullptr;
  BSK_CUDA(cudaMalloc((void **)&nk, cap * 8));
  BSK_CUDA(cudaMalloc((void **)&nf, cap * 8));
  if (rm->n_hist) {
    BSK_CUDA(cudaMemcpyAsync(nk, rm->hist_keys, rm->n_hist * 8, cudaMemcpyDeviceToDevice, s));
    BSK_CUDA(cudaMemcpyAsync(nf, rm->hist_fp, rm->n_hist * 8, cudaMemcpyDeviceToDevice, s));
    BSK_CUDA(cudaStreamSynchronize(s));
  }
  if (rm->hist_keys) cudaFree(rm->hist_keys);
  if (rm->hist_fp) cudaFree(rm->hist_fp);
  rm->hist_keys = nk;
  rm->hist_fp = nf;
  rm->hist_cap = cap;
}

static void table_reserve(Engine::RmdupState *rm, u64 total, cudaStream_t s, u64 &launches) {
  if (rm->cap && !rm->dirty && total * 3 <= rm->cap * 2) return;  // load factor <= 2/3
  u64 cap = 1u << 12;
  while (cap < total * 2) cap *= 2;
  if (cap > rm->alloc_cap) {  // device memory is kept across bsk_reset / partitions: only growth reallocates
    if (rm->table) cudaFree(rm->table);
    rm->table = nullptr;
    BSK_CUDA(cudaMalloc((void **)&rm->table, (cap + 1) * sizeof(TableSlot)));
    rm->alloc_cap = cap;
  }
  BSK_CUDA(cudaMemsetAsync(rm->table, 0xff, (cap + 1) * sizeof(TableSlot), s));
  rm->cap = cap;
  rm->dirty = false;
  if (rm->n_hist) {
    BSK_LAUNCH_FLAT(k_table_insert, (u32)((rm->n_hist + 255) / 256), 256, 0, s, rm->hist_keys, rm->n_hist, (u64)0, rm->table,
                    cap);
    launches++;
  }
}

void rmdup_state_free(Engine::RmdupState *rm) {
  if (!rm) return;
  if (rm->table) cudaFree(rm->table);
  if (rm->hist_keys) cudaFree(rm->hist_keys);
  if (rm->hist_fp) cudaFree(rm->hist_fp);
  delete rm;
}

void rmdup_state_reset(Engine::RmdupState *rm) {
  if (!rm) return;
  rm->n_hist = 0;
  rm->dup_seqs.clear();
  rm->id_text.clear();
  rm->id_off.clear();
  rm->dup_pairs.clear();
  rm->dup_num.clear();
  rm->dirty = true;  // the table is cleared (not freed) before its next use
  rm->block_ready = false;
}

int Engine::rmdup_hash_block() {
  if (!rm_) rm_ = new RmdupState();
  const size_t R = (size_t)n_rec_ + 1;
  SubjectViews sv;
  if (o_.BySeq) {
    sv = SubjectViews{views_.seqb, views_.seq_off, views_.seq_len, views_.seqb == in_ ? (u64)n_ : (u64)seq_space_};
  } else if (o_.ByName) {
    sv = SubjectViews{in_, ra_.head_off, ra_.head_len, (u64)n_};
  } else {
    u32 *ids = b_id_.get<u32>(R * 2);
    k::id_desc(views_, o_.IDNCBI ? 1 : 0, ids, ids + R, nullptr, nullptr, stream);
    launches_++;
    sv = SubjectViews{in_, ids, ids + R, (u64)n_};
  }
  rm_->sv_base = sv.base;
  rm_->sv_off = sv.off;
  rm_->sv_len = sv.len;
  rm_->sv_limit = sv.limit;
  u64 *keys = b_op1_.get<u64>(R);
  u64 *fps = b_op2_.get<u64>(R);
  if (n_rec_) {
    main_begin();
    BSK_LAUNCH_FLAT(k_rmdup_hash, (n_rec_ + 127) / 128, 128, 0, stream, sv, n_rec_, o_.IgnoreCase ? 1 : 0, keys, fps);
    main_end();
    launches_++;
  }
  return BSK_OK;
}

int Engine::rmdup_resolve_block(BlockOut &bo) {
  RmdupState *rm = rm_;
  const u64 g_base = rm->n_hist;
  u64 *keys = b_op1_.as<u64>(), *fps = b_op2_.as<u64>();
  SubjectViews sv{rm->sv_base, rm->sv_off, rm->sv_len, rm->sv_limit};
  u8 *keep = b_keep_.get<u8>((size_t)n_rec_ + 1);
  const bool want_dup_seqs = !o_.DupSeqsFile.empty(), want_dup_num = !o_.DupNumFile.empty();
  u64 *first = want_dup_num ? b_op5_.get<u64>((size_t)n_rec_ + 1) : nullptr;
  if (n_rec_) {
    table_reserve(rm, g_base + n_rec_, stream, launches_);
    BSK_LAUNCH_FLAT(k_table_insert, (n_rec_ + 255) / 256, 256, 0, stream, keys, (u64)n_rec_, g_base, rm->table, rm->cap);
    BSK_LAUNCH_FLAT(k_rmdup_resolve, (n_rec_ + 255) / 256, 256, 0, stream, sv, n_rec_, o_.IgnoreCase ? 1 : 0, keys, fps,
                    g_base, rm->hist_fp, rm->table, rm->cap, keep, first, d_status_);
    launches_ += 2;
    fetch_status();
    if (h_status_->counters[4]) {
      BSK_LAUNCH_FLAT(k_rmdup_fixup, 1, 1, 0, stream, sv, n_rec_, o_.IgnoreCase ? 1 : 0, keys, fps, g_base, rm->hist_keys,
                      rm->hist_fp, keep, first);
      launches_++;
    }
    hist_reserve(rm, g_base + n_rec_, stream);
    BSK_CUDA(cudaMemcpyAsync(rm->hist_keys + g_base, keys, (size_t)n_rec_ * 8, cudaMemcpyDeviceToDevice, stream));
    BSK_CUDA(cudaMemcpyAsync(rm->hist_fp + g_base, fps, (size_t)n_rec_ * 8, cudaMemcpyDeviceToDevice, stream));
    rm->n_hist = g_base + n_rec_;
  }
  if (n_rec_ && (want_dup_seqs || want_dup_num)) {
    int rc = rmdup_side_outputs(keep, first, g_base);
    if (rc != BSK_OK) return rc;
  }
  // Record.Format(LineWidth) minus the final '\n' (rmdup.go:214-215) + FileStore's '\n'
  EmitCfg cfg;
  cfg.marker = fastq_ ? '@' : '>';
  cfg.print_name = 1;
  cfg.print_seq = 1;
  cfg.print_qual = fastq_;
  cfg.plus_line = fastq_;
  cfg.reverse = 0;
  cfg.width = fastq_ ? 0 : (o_.LineWidth > 0 ? (u32)o_.LineWidth : 0);
  views_.name_off = ra_.head_off;
  views_.name_len = ra_.head_len;
  int rc = emit_records(cfg, n_rec_ ? keep : nullptr, nullptr, bo);
  if (rc == BSK_OK) rmdup_removed += n_rec_ - bo.n_elem;
  return rc;
}

// rmdup -d / -D (RmDupCheck.Call bigseqkit-lib/rmdup.go:180-239): the removed records as Record.Format(LineWidth) and,
// per record, {group's first ordinal, ID}.  Both are accumulated on the host like the reference's this.dups /
// this.data until bsk_rmdup_dup_seqs / bsk_rmdup_dup_num fetch them (After, rmdup.go:245-275).  Runs before the
// survivors are emitted: it borrows the output buffers.
int Engine::rmdup_side_outputs(const u8 *keep, const u64 *first, u64 g_base) {
  RmdupState *rm = rm_;
  const size_t R = (size_t)n_rec_ + 1;
  const int saved_elem = want_elem_off;
  want_elem_off = 0;
  int rc = BSK_OK;
  if (!o_.DupSeqsFile.empty()) {
    u8 *drop = b_op4_.get<u8>(R);
    BSK_LAUNCH_FLAT(k_rmdup_drop_mask, (n_rec_ + 255) / 256, 256, 0, stream, keep, n_rec_, drop);
    launches_++;
    EmitCfg cfg;
    cfg.marker = fastq_ ? '@' : '>';
    cfg.print_name = 1;
    cfg.print_seq = 1;
    cfg.print_qual = fastq_;
    cfg.plus_line = fastq_;
    cfg.reverse = 0;
    cfg.width = fastq_ ? 0 : (o_.LineWidth > 0 ? (u32)o_.LineWidth : 0);
    views_.name_off = ra_.head_off;
    views_.name_len = ra_.head_len;
    BlockOut side;
    rc = emit_records(cfg, drop, nullptr, side);
    if (rc == BSK_OK && side.n) {
      const size_t at = rm->dup_seqs.size();
      rm->dup_seqs.resize(at + side.n);
      BSK_CUDA(cudaMemcpyAsync(&rm->dup_seqs[at], side.d_data, side.n, cudaMemcpyDeviceToHost, stream));
      BSK_CUDA(cudaStreamSynchronize(stream));
    }
  }
  if (rc == BSK_OK && !o_.DupNumFile.empty()) {
    u32 *ids = b_op6_.get<u32>(R * 2);
    views_.name_off = ra_.head_off;
    views_.name_len = ra_.head_len;
    k::id_desc(views_, o_.IDNCBI ? 1 : 0, ids, ids + R, nullptr, nullptr, stream);
    launches_++;
    views_.name_off = ids;
    views_.name_len = ids + R;
    EmitCfg cfg{};
    cfg.print_name = 1;  // one "ID\n" per record
    BlockOut side;
    rc = emit_records(cfg, nullptr, nullptr, side);
    views_.name_off = ra_.head_off;
    views_.name_len = ra_.head_len;
    if (rc == BSK_OK) {
      const size_t at = rm->id_text.size();
      rm->id_text.resize(at + side.n);
      std::vector<u8> hk(n_rec_);
      std::vector<u64> hf(n_rec_);
      if (side.n) BSK_CUDA(cudaMemcpyAsync(&rm->id_text[at], side.d_data, side.n, cudaMemcpyDeviceToHost, stream));
      BSK_CUDA(cudaMemcpyAsync(hk.data(), keep, n_rec_, cudaMemcpyDeviceToHost, stream));
      BSK_CUDA(cudaMemcpyAsync(hf.data(), first, (size_t)n_rec_ * 8, cudaMemcpyDeviceToHost, stream));
      BSK_CUDA(cudaStreamSynchronize(stream));
      size_t pos = at;
      for (u32 r = 0; r < n_rec_; r++) {
        rm->id_off.push_back(pos);
        const void *nl = memchr(rm->id_text.data() + pos, '\n', rm->id_text.size() - pos);
        pos = nl ? (size_t)((const char *)nl - rm->id_text.data()) + 1 : rm->id_text.size();
        if (hk[r] == 0) rm->dup_pairs.emplace_back(hf[r], g_base + r);
      }
    }
  }
  want_elem_off = saved_elem;
  return rc;
}

int Engine::rmdup_dup_seqs(const char **data, size_t *n) {
  if (op_ != OP_RMDUP) { err = "bsk_rmdup_dup_seqs: ctx is not an RmDup operator"; return BSK_ERR_STATE; }
  *data = rm_ ? rm_->dup_seqs.data() : "";
  *n = rm_ ? rm_->dup_seqs.size() : 0;
  return BSK_OK;
}

// rows "count\tid1, id2, ...\n" (rmdup.go:230-234), one per subject with more than one member, in the order of the
// groups' first members; ids in input order
int Engine::rmdup_dup_num(const char **data, size_t *n) {
  if (op_ != OP_RMDUP) { err = "bsk_rmdup_dup_num: ctx is not an RmDup operator"; return BSK_ERR_STATE; }
  *data = "";
  *n = 0;
  if (!rm_) return BSK_OK;
  RmdupState *rm = rm_;
  std::stable_sort(rm->dup_pairs.begin(), rm->dup_pairs.end(),
                   [](const std::pair<u64, u64> &a, const std::pair<u64, u64> &b) { return a.first < b.first; });
  auto id_of = [&](u64 g, const char *&p, size_t &l) {
    p = rm->id_text.data() + rm->id_off[g];
    const size_t end = g + 1 < rm->id_off.size() ? rm->id_off[g + 1] : rm->id_text.size();
    l = end - rm->id_off[g] - 1;
  };
  std::string &out = rm->dup_num;
  out.clear();
  for (size_t i = 0; i < rm->dup_pairs.size();) {
    size_t j = i;
    while (j < rm->dup_pairs.size() && rm->dup_pairs[j].first == rm->dup_pairs[i].first) j++;
    if (rm->dup_pairs[i].first >= rm->id_off.size()) { err = "bsk_rmdup_dup_num: first member of a group is outside this ctx"; return BSK_ERR_STATE; }
    out += std::to_string(j - i + 1);
    out += '\t';
    const char *p;
    size_t l;
    id_of(rm->dup_pairs[i].first, p, l);
    out.append(p, l);
    for (size_t k = i; k < j; k++) {
      id_of(rm->dup_pairs[k].second, p, l);
      out += ", ";
      out.append(p, l);
    }
    out += '\n';
    i = j;
  }
  *data = out.data();
  *n = out.size();
  return BSK_OK;
}

int Engine::op_rmdup(BlockOut &bo, bool prepare_only) {
  int rc = check_errors();
  if (rc != BSK_OK) return rc;
  if (first_block_ && !union_) {  // a new partition: forget the previous one (unless the partitions are one Union)
    if (rm_) rmdup_state_reset(rm_);
    rmdup_removed = 0;
  }
  rc = rmdup_hash_block();
  if (rc != BSK_OK) return rc;
  return rmdup_finish(bo, prepare_only);
}

// keys / fingerprints of the block are in b_op1_ / b_op2_ and rm_->sv_* describes the subjects: resolve + emit
int Engine::rmdup_finish(BlockOut &bo, bool prepare_only) {
  if (!prepare_only) return rmdup_resolve_block(bo);
  // RmDupPrepare: every record formatted, keys kept for bsk_rmdup_keys
  RmdupState *rm = rm_;
  if (n_rec_) {
    hist_reserve(rm, rm->n_hist + n_rec_, stream);
    BSK_CUDA(cudaMemcpyAsync(rm->hist_keys + rm->n_hist, b_op1_.p, (size_t)n_rec_ * 8, cudaMemcpyDeviceToDevice, stream));
    BSK_CUDA(cudaMemcpyAsync(rm->hist_fp + rm->n_hist, b_op2_.p, (size_t)n_rec_ * 8, cudaMemcpyDeviceToDevice, stream));
    rm->n_hist += n_rec_;
  }
  EmitCfg cfg;
  cfg.marker = fastq_ ? '@' : '>';
  cfg.print_name = 1;
  cfg.print_seq = 1;
  cfg.print_qual = fastq_;
  cfg.plus_line = fastq_;
  cfg.reverse = 0;
  cfg.width = fastq_ ? 0 : (o_.LineWidth > 0 ? (u32)o_.LineWidth : 0);
  views_.name_off = ra_.head_off;
  views_.name_len = ra_.head_len;
  return emit_records(cfg, nullptr, nullptr, bo);
}

// Short 4-line FASTQ records: index + parse + hash in one streaming kernel (k_rmdup_tile.cu), then the same
// resolve / emit as the general path.  kFusedFallback when the block is outside that grammar.
int Engine::op_rmdup_tile(const u8 *d_in, u32 n, BlockOut &bo, bool prepare_only) {
  if (n == 0 || !fused_ok_ || o_.IgnoreCase || getenv("BSK_NO_RMDUP_TILE")) return kFusedFallback;
  bool fastq = false, ok = false;
  const int saved_alpha = alphabet_;
  const bool saved_known = alphabet_known_;
  if (!alphabet_known_ || first_block_) {
    int rc = first_record_alphabet(d_in, n, fastq, ok);
    if (rc != BSK_OK) return rc;
    if (!ok || !fastq) { alphabet_ = saved_alpha; alphabet_known_ = saved_known; return kFusedFallback; }
  } else {
    fastq = part_fastq_;
    if (!fastq) return kFusedFallback;
  }
  if (!n_sm_) {
    cudaDeviceProp prop;
    BSK_CUDA(cudaGetDeviceProperties(&prop, device_ >= 0 ? device_ : 0));
    n_sm_ = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 1;
  }
  reset_status();
  BSK_CUDA(cudaEventRecord(ev_[1], stream));
  const u32 n_tiles = k::rmdup_tile_tiles(n);
  void *slots = b_op3_.get<u8>((size_t)n_tiles * k::rmdup_tile_slot_stride() * k::rmdup_tile_slot_bytes());
  u32 *tile_cnt = b_tile_cnt_.get<u32>((size_t)n_tiles + 1);
  BSK_CUDA(cudaMemsetAsync(tile_cnt, 0, ((size_t)n_tiles + 1) * 4, stream));
  const int subject = o_.BySeq ? 0 : (o_.ByName ? 1 : 2);
  main_begin();
  k::rmdup_tile(d_in, n, slots, tile_cnt, d_status_, subject, n_sm_, stream);
  main_end();
  launches_++;
  u64 *tile_base = b_tile_base_.get<u64>((size_t)n_tiles + 1);
  prim::excl_scan_u32_to_u64(tile_cnt, tile_base, (size_t)n_tiles + 1, b_tmp_, stream);
  u8 *hs = h_small_.as<u8>();
  prim::copy_small(hs, tile_base + n_tiles, 8, stream);
  fetch_status();  // synchronises the stream
  if (h_status_->counters[0] || (o_.IDNCBI && subject == 2)) {
    alphabet_ = saved_alpha;
    alphabet_known_ = saved_known;
    main_timed_ = false;
    timings.main_launches--;
    return kFusedFallback;
  }
  u64 nrec;
  memcpy(&nrec, hs, 8);
  // the block state the general path would have left behind (no line index on this path)
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
  u64 *keys = b_op1_.get<u64>(R);
  u64 *fps = b_op2_.get<u64>(R);
  u32 *ids = b_id_.get<u32>(R * 2);
  k::rmdup_tile_compact(slots, tile_cnt, tile_base, n_tiles, keys, fps, ra_, ids + R, stream);
  launches_++;
  seq_space_ = qual_space_ = n;
  set_views_default();
  bo.n_rec = n_rec_;
  if (n_rec_) any_record_ = true;
  if (first_block_ && !union_) {  // a new partition: forget the previous one (unless the partitions are one Union)
    if (rm_) rmdup_state_reset(rm_);
    rmdup_removed = 0;
  }
  if (!rm_) rm_ = new RmdupState();
  rm_->sv_base = d_in;
  rm_->sv_limit = n;
  rm_->sv_off = subject == 0 ? ra_.seq_off : ra_.head_off;
  rm_->sv_len = subject == 0 ? ra_.seq_len : (subject == 1 ? ra_.head_len : ids + R);
  timings.fused_blocks++;
  contig_known_ = true;  // the tile kernel accepted the block: every record is printed exactly as it stands in the input
  const int frc = rmdup_finish(bo, prepare_only);
  contig_known_ = false;
  return frc;
}

int Engine::rmdup_keys(const int64_t **keys, size_t *n) {
  if (op_ != OP_RMDUP && op_ != OP_RMDUP_PREPARE) { err = "bsk_rmdup_keys: ctx is not an RmDup operator"; return BSK_ERR_STATE; }
  if (device_ >= 0) BSK_CUDA(cudaSetDevice(device_));
  const u64 cnt = rm_ ? rm_->n_hist : 0;
  keys_host_.resize(cnt);
  if (cnt) BSK_CUDA(cudaMemcpy(keys_host_.data(), rm_->hist_keys, cnt * 8, cudaMemcpyDeviceToHost));
  *keys = keys_host_.data();
  *n = cnt;
  return BSK_OK;
}

// multi-GPU step 1: index + hash the local shard; keys / second hashes stay in b_op1_ / b_op2_
int Engine::rmdup_prepare_local(const void *d_in, size_t n, u64 *n_records) {
  if (op_ != OP_RMDUP) { err = "bsk_rmdup_prepare_device: ctx is not an RmDup operator"; return BSK_ERR_STATE; }
  if (n >= kMaxBlockBytes) { err = "bsk_rmdup_prepare_device: shard must be smaller than 4 GiB - 1 MiB"; return BSK_ERR_ARG; }
  if (((uintptr_t)d_in & 15) != 0) { err = "bsk_rmdup_prepare_device: device pointer must be 16-byte aligned"; return BSK_ERR_ARG; }
  if (!o_.DupNumFile.empty()) {  // the first member of a group may live on another rank
    err = "-D/--dup-num-file needs the whole input in one ctx (bsk_run_buffer / bsk_run_file), not the sharded path";
    return BSK_ERR_UNSUPPORTED;
  }
  if (device_ >= 0) BSK_CUDA(cudaSetDevice(device_));
  launches_ = 0;
  timings = bsk_timings{};
  alphabet_ = o_.alphabet;
  alphabet_known_ = false;
  first_block_ = true;
  main_timed_ = false;
  if (!rm_) rm_ = new RmdupState();
  rmdup_state_reset(rm_);
  rmdup_removed = 0;
  BSK_CUDA(cudaEventRecord(ev_[0], stream));
  int rc = prepare_block(static_cast<const u8 *>(d_in), (u32)n);
  if (rc != BSK_OK) return rc;
  rc = resolve_alphabet();
  if (rc != BSK_OK) return rc;
  BSK_CUDA(cudaEventRecord(ev_[1], stream));
  rc = check_errors();
  if (rc != BSK_OK) return rc;
  if (n_records) *n_records = n_rec_;
  rc = rmdup_hash_block();
  if (rc != BSK_OK) return rc;
  BSK_CUDA(cudaEventRecord(ev_[4], stream));
  BSK_CUDA(cudaStreamSynchronize(stream));
  accumulate_timings();
  timings.kernel_launches = launches_;
  timings.in_bytes = n;
  rm_->block_ready = true;
  first_block_ = false;
  return BSK_OK;
}

void Engine::rmdup_export_fp(u64 *d_fp) {
  if (!n_rec_) return;
  BSK_LAUNCH_FLAT(k_interleave_fp, (n_rec_ + 255) / 256, 256, 0, stream, b_op1_.as<u64>(), b_op2_.as<u64>(), (u64)n_rec_, d_fp);
  launches_++;
}

int Engine::rmdup_prepare_device(const void *d_in, size_t n, void *d_fp, size_t fp_cap, u64 *n_records) {
  u64 nr = 0;
  int rc = rmdup_prepare_local(d_in, n, &nr);
  if (n_records) *n_records = nr;
  if (rc != BSK_OK) return rc;
  if ((size_t)nr > fp_cap) { rm_->block_ready = false; err = "bsk_rmdup_prepare_device: fingerprint buffer too small"; return BSK_ERR_ARG; }
  rmdup_export_fp(static_cast<u64 *>(d_fp));
  BSK_CUDA(cudaStreamSynchronize(stream));
  timings.kernel_launches = launches_;
  return BSK_OK;
}

// multi-GPU step 3: d_all_fp holds the fingerprints of the n_before records that precede this shard
// in global input order (what the all-gather delivered); survivors of the local shard are emitted.
int Engine::rmdup_resolve_device(const void *d_all_fp, u64 n_before, bsk_out *out) {
  memset(out, 0, sizeof *out);
  if (!rm_ || !rm_->block_ready) { err = "bsk_rmdup_resolve_device: call bsk_rmdup_prepare_device first"; return BSK_ERR_STATE; }
  if (device_ >= 0) BSK_CUDA(cudaSetDevice(device_));
  RmdupState *rm = rm_;
  rm->block_ready = false;
  main_timed_ = false;
  BSK_CUDA(cudaEventRecord(ev_[0], stream));
  BSK_CUDA(cudaEventRecord(ev_[1], stream));
  if (n_before) {
    hist_reserve(rm, n_before, stream);
    BSK_LAUNCH_FLAT(k_deinterleave_fp, (u32)((n_before + 255) / 256), 256, 0, stream, static_cast<const u64 *>(d_all_fp),
                    n_before, rm->hist_keys, rm->hist_fp);
    launches_++;
    rm->n_hist = n_before;
    table_reserve(rm, n_before + n_rec_, stream, launches_);
  }
  BlockOut bo;
  int rc = rmdup_resolve_block(bo);
  BSK_CUDA(cudaEventRecord(ev_[4], stream));
  BSK_CUDA(cudaStreamSynchronize(stream));
  if (rc != BSK_OK) return rc;
  accumulate_timings();
  timings.kernel_launches = launches_;
  timings.out_bytes = bo.n;
  out->data = bo.d_data;
  out->n = bo.n;
  out->elem_off = bo.d_elem_off;
  out->n_elem = bo.n_elem;
  out->n_records = n_rec_;
  return BSK_OK;
}

}  // namespace bsk
