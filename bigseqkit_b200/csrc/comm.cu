// comm.cu -- the two exchange steps of the path, behind the C ABI:
//
//   stats   StatsReduce over all partitions   bigseqkit/stats.go:91 (Reduce), bigseqkit-lib/stats.go:128-137
//   rmdup   GroupByKey over all partitions    bigseqkit/rmdup.go:97, then RmDupCheck bigseqkit-lib/rmdup.go:118-242
//   ordered merged output (FileStore's MPI token ring, bigseqkit-lib/helper.go:399-431) -> one all-gather of sizes
//
// Two forms with the same semantics (sum for stats; first occurrence in global input order for rmdup, Q4):
//   * one process per GPU: a NCCL communicator bound to the ctx (bsk_comm_init), collectives on the ctx stream;
//   * several ctxs of ONE process (bsk_reduce / bsk_rmdup_union): plain device copies, no NCCL.
// NCCL is loaded lazily (dlopen "libnccl.so.2"): libbsk.so itself only needs the CUDA runtime, and a process that
// already holds a NCCL (e.g. PyTorch's) shares it.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "engine.h"
#include "op_state.h"

#ifndef BSK_EMU
#include <dlfcn.h>
#endif

namespace bsk {

// ---- minimal NCCL surface (matches nccl.h 2.x; the types below are ABI-stable since 2.0)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclChar = 0, ncclUint64 = 5 };
enum { ncclSum = 0 };

struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int *) = nullptr;
  std::string err;
};

static NcclApi *nccl_api(std::string &err) {
  static NcclApi api;
  static bool tried = false;
  if (tried) {
    if (!api.lib) err = api.err;
    return api.lib ? &api : nullptr;
  }
  tried = true;
#ifdef BSK_EMU
  api.err = "NCCL is not available in the host emulator";
#else
  const char *names[4] = {getenv("BSK_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
  // a NCCL already mapped into the process (PyTorch's) wins: same soname
  api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  for (int i = 0; !api.lib && i < 3; i++)
    if (names[i] && names[i][0]) api.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!api.lib) {
    api.err = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found") +
              "; set BSK_NCCL_LIB to its path";
  } else {
    bool ok = true;
    auto sym = [&](const char *n) { void *p = dlsym(api.lib, n); if (!p) ok = false; return p; };
    api.GetUniqueId = (int (*)(ncclUniqueId *))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))sym("ncclCommInitRank");
    api.CommDestroy = (int (*)(ncclComm_t))sym("ncclCommDestroy");
    api.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))sym("ncclAllReduce");
    api.AllGather = (int (*)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t))sym("ncclAllGather");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
    api.GetVersion = (int (*)(int *))sym("ncclGetVersion");
    if (!ok) { api.err = "libnccl.so.2 lacks a required symbol"; api.lib = nullptr; }
  }
#endif
  if (!api.lib) err = api.err;
  return api.lib ? &api : nullptr;
}

struct Engine::CommState {
  NcclApi *api = nullptr;
  ncclComm_t comm = nullptr;
  int n = 1, rank = 0;
};

#define BSK_NCCL(call)                                                                              \
  do {                                                                                              \
    const int r_ = (call);                                                                          \
    if (r_ != ncclSuccess)                                                                          \
      throw ::bsk::CudaError(std::string(#call) + ": " + comm_->api->GetErrorString(r_));           \
  } while (0)

int comm_unique_id(uint8_t *id, std::string &err) {
  NcclApi *api = nccl_api(err);
  if (!api) return BSK_ERR_UNSUPPORTED;
  ncclUniqueId u;
  const int r = api->GetUniqueId(&u);
  if (r != ncclSuccess) { err = std::string("ncclGetUniqueId: ") + api->GetErrorString(r); return BSK_ERR_CUDA; }
  memcpy(id, u.internal, sizeof u.internal);
  return BSK_OK;
}

int Engine::comm_init(const uint8_t *id, int n_ranks, int rank) {
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) { err = "bsk_comm_init: bad rank / n_ranks"; return BSK_ERR_ARG; }
  comm_free();
  NcclApi *api = nccl_api(err);
  if (!api) return BSK_ERR_UNSUPPORTED;
  if (device_ >= 0) BSK_CUDA(cudaSetDevice(device_));
  comm_ = new CommState();
  comm_->api = api;
  comm_->n = n_ranks;
  comm_->rank = rank;
  ncclUniqueId u;
  memcpy(u.internal, id, sizeof u.internal);
  const int r = api->CommInitRank(&comm_->comm, n_ranks, u, rank);
  if (r != ncclSuccess) {
    err = std::string("ncclCommInitRank: ") + api->GetErrorString(r);
    delete comm_;
    comm_ = nullptr;
    return BSK_ERR_CUDA;
  }
  return BSK_OK;
}

void Engine::comm_free() {
  if (!comm_) return;
  if (comm_->comm) {
    if (device_ >= 0) cudaSetDevice(device_);
    cudaStreamSynchronize(stream);
    comm_->api->CommDestroy(comm_->comm);
  }
  delete comm_;
  comm_ = nullptr;
}

int Engine::comm_rank(int *rank, int *n_ranks) const {
  if (rank) *rank = comm_ ? comm_->rank : 0;
  if (n_ranks) *n_ranks = comm_ ? comm_->n : 1;
  return BSK_OK;
}

// every rank contributes n_words u64; h_all receives n_ranks * n_words (rank order).  Synchronises the ctx stream.
void Engine::comm_gather_words(const u64 *h_mine, size_t n_words, u64 *h_all) {
  const size_t n = (size_t)comm_->n;
  u64 *d = b_comm_small_.get<u64>((n + 1) * n_words);
  BSK_CUDA(cudaMemcpyAsync(d, h_mine, n_words * 8, cudaMemcpyHostToDevice, stream));
  BSK_NCCL(comm_->api->AllGather(d, d + n_words, n_words * 8, ncclChar, comm_->comm, stream));
  BSK_CUDA(cudaMemcpyAsync(h_all, d + n_words, n * n_words * 8, cudaMemcpyDeviceToHost, stream));
  BSK_CUDA(cudaStreamSynchronize(stream));
}

// where this rank's bytes go in the merged output: exclusive prefix of the sizes in rank order
int Engine::output_offsets(u64 n_local, u64 *offset, u64 *total) {
  u64 off = 0, tot = n_local;
  if (comm_ && comm_->n > 1) {
    if (device_ >= 0) BSK_CUDA(cudaSetDevice(device_));
    std::vector<u64> all((size_t)comm_->n);
    comm_gather_words(&n_local, 1, all.data());
    tot = 0;
    for (int r = 0; r < comm_->n; r++) {
      if (r < comm_->rank) off += all[(size_t)r];
      tot += all[(size_t)r];
    }
  }
  if (offset) *offset = off;
  if (total) *total = tot;
  return BSK_OK;
}

// ------------------------------------------------------------------ stats
// Lengths below kDenseBins travel as a dense u64 histogram together with the scalar sums through ONE all-reduce;
// {records, long-length pairs, type} of every rank through one small all-gather; lengths >= kDenseBins (contigs)
// as padded (length, count) pairs through a second all-gather only when some rank has any.
static const size_t kDenseBins = 65536;

int Engine::stats_allreduce() {
  if (op_ != OP_STATS) { err = "bsk_stats_allreduce: ctx is not a Stats operator"; return BSK_ERR_STATE; }
  if (!comm_) { err = "bsk_stats_allreduce: no communicator (bsk_comm_init)"; return BSK_ERR_STATE; }
  if (comm_->n == 1) return BSK_OK;
  if (device_ >= 0) BSK_CUDA(cudaSetDevice(device_));
  const size_t n = (size_t)comm_->n;
  std::vector<u64> long_len, long_cnt;
  u64 num = 0;
  for (auto &kv : hist_) {
    num += kv.second;
    if (kv.first >= kDenseBins) { long_len.push_back(kv.first); long_cnt.push_back(kv.second); }
  }
  // rank info: records, long pairs, type
  u64 mine[4] = {num, (u64)long_len.size(), 0, 0};
  memcpy(&mine[2], stats_type_.c_str(), std::min<size_t>(stats_type_.size(), 15));
  std::vector<u64> info(n * 4);
  comm_gather_words(mine, 4, info.data());
  // dense histogram + scalars
  u64 *d = b_comm_.get<u64>(kDenseBins + 8);
  u64 over = 0;
  int rc = stats_dense_device(d, kDenseBins, &over);  // synchronises
  if (rc != BSK_OK) return rc;
  u64 *hs = h_small_.as<u64>();
  hs[0] = q20_; hs[1] = q30_; hs[2] = gap_; hs[3] = 0; hs[4] = 0; hs[5] = 0; hs[6] = 0; hs[7] = 0;
  BSK_CUDA(cudaMemcpyAsync(d + kDenseBins, hs, 64, cudaMemcpyHostToDevice, stream));
  BSK_NCCL(comm_->api->AllReduce(d, d, kDenseBins + 8, ncclUint64, ncclSum, comm_->comm, stream));
  launches_++;
  std::vector<u64> dense(kDenseBins + 8);
  BSK_CUDA(cudaMemcpyAsync(dense.data(), d, (kDenseBins + 8) * 8, cudaMemcpyDeviceToHost, stream));
  u64 max_long = 0;
  for (size_t r = 0; r < n; r++) max_long = std::max(max_long, info[r * 4 + 1]);
  std::vector<u64> longs;
  if (max_long) {
    std::vector<u64> pad(2 * max_long, 0);
    for (size_t i = 0; i < long_len.size(); i++) { pad[2 * i] = long_len[i]; pad[2 * i + 1] = long_cnt[i]; }
    longs.resize(n * 2 * max_long);
    BSK_CUDA(cudaStreamSynchronize(stream));
    comm_gather_words(pad.data(), 2 * max_long, longs.data());
  }
  BSK_CUDA(cudaStreamSynchronize(stream));
  // every rank now holds the global totals
  hist_.clear();
  for (size_t l = 0; l < kDenseBins; l++)
    if (dense[l]) hist_[l] = dense[l];
  for (size_t r = 0; r < n; r++)
    for (u64 i = 0; i < info[r * 4 + 1]; i++) hist_[longs[(r * max_long + i) * 2]] += longs[(r * max_long + i) * 2 + 1];
  q20_ = dense[kDenseBins];
  q30_ = dense[kDenseBins + 1];
  gap_ = dense[kDenseBins + 2];
  // the type column comes from the first rank that saw a record (reference: partition 0's first record)
  for (size_t r = 0; r < n; r++)
    if (info[r * 4]) {
      char t[17] = {0};
      memcpy(t, &info[r * 4 + 2], 16);
      stats_type_ = t;
      stats_type_set_ = true;
      break;
    }
  return BSK_OK;
}

// several ctxs of one process: every ctx ends with the sum of all (StatsReduce folded over the partitions)
int stats_reduce_local(Engine **e, int n, std::string &err) {
  if (n < 1) return BSK_OK;
  for (int i = 1; i < n; i++) {
    const int rc = e[0]->stats_merge_from(*e[i]);
    if (rc != BSK_OK) { err = e[0]->err; return rc; }
  }
  for (int i = 1; i < n; i++) {
    e[i]->stats_clear_totals();
    const int rc = e[i]->stats_merge_from(*e[0]);
    if (rc != BSK_OK) { err = e[i]->err; return rc; }
  }
  return BSK_OK;
}

// ------------------------------------------------------------------ rmdup
// one process per GPU: hash the local shard, all-gather the 16-byte fingerprints (padded to the largest shard),
// keep the ones of earlier ranks, resolve.  A rank whose shard fails to parse still takes part in the count
// exchange (count = ~0) so that every rank returns an error instead of hanging in the collective.
int Engine::rmdup_sharded(const void *d_in, size_t n, bsk_out *out) {
  memset(out, 0, sizeof *out);
  if (op_ != OP_RMDUP) { err = "bsk_rmdup_sharded: ctx is not an RmDup operator"; return BSK_ERR_STATE; }
  if (!comm_ || comm_->n == 1) return run_device(d_in, n, comm_ ? comm_->rank : 0, out);
  if (device_ >= 0) BSK_CUDA(cudaSetDevice(device_));
  const size_t nr = (size_t)comm_->n;
  u64 n_local = 0;
  int rc = rmdup_prepare_local(d_in, n, &n_local);
  std::string local_err = err;
  u64 mine = rc == BSK_OK ? n_local : ~0ull;
  std::vector<u64> counts(nr);
  comm_gather_words(&mine, 1, counts.data());
  u64 mx = 1, n_before = 0;
  for (size_t r = 0; r < nr; r++) {
    if (counts[r] == ~0ull) {
      if (rc == BSK_OK) { err = "bsk_rmdup_sharded: rank " + std::to_string(r) + " failed on its shard"; rc = BSK_ERR_DATA; }
      else err = local_err;
      return rc;
    }
    mx = std::max(mx, counts[r]);
    if ((int)r < comm_->rank) n_before += counts[r];
  }
  u64 *send = b_fp_.get<u64>(2 * mx);
  u64 *recv = b_comm_.get<u64>(2 * mx * nr);
  rmdup_export_fp(send);
  BSK_NCCL(comm_->api->AllGather(send, recv, mx * 16, ncclChar, comm_->comm, stream));
  launches_++;
  // fingerprints of the records that precede this shard, contiguous
  u64 *before = b_fp_before_.get<u64>(2 * std::max<u64>(n_before, 1));
  u64 pos = 0;
  for (int r = 0; r < comm_->rank; r++) {
    if (counts[(size_t)r])
      BSK_CUDA(cudaMemcpyAsync(before + 2 * pos, recv + 2 * mx * (u64)r, counts[(size_t)r] * 16, cudaMemcpyDeviceToDevice, stream));
    pos += counts[(size_t)r];
  }
  return rmdup_resolve_device(before, n_before, out);
}

// several ctxs of one process (any devices), ctx order == input order
int rmdup_union_local(Engine **e, int n, const void *const *d_in, const size_t *nbytes, bsk_out *outs, std::string &err) {
  std::vector<u64> counts((size_t)n, 0);
  for (int i = 0; i < n; i++) {
    memset(&outs[i], 0, sizeof outs[i]);
    const int rc = e[i]->rmdup_prepare_local(d_in[i], nbytes[i], &counts[(size_t)i]);
    if (rc != BSK_OK) { err = e[i]->err; return rc; }
    e[i]->rmdup_export_own_fp();
  }
  for (int i = 0; i < n; i++) {
    const int rc = e[i]->rmdup_resolve_after(e, i, counts.data(), &outs[i]);
    if (rc != BSK_OK) { err = e[i]->err; return rc; }
  }
  return BSK_OK;
}

void Engine::rmdup_export_own_fp() {
  if (device_ >= 0) BSK_CUDA(cudaSetDevice(device_));
  u64 *send = b_fp_.get<u64>(2 * std::max<u64>(n_rec_, 1));
  rmdup_export_fp(send);
  BSK_CUDA(cudaStreamSynchronize(stream));
}

int Engine::rmdup_resolve_after(Engine **e, int self, const u64 *counts, bsk_out *out) {
  if (device_ >= 0) BSK_CUDA(cudaSetDevice(device_));
  u64 n_before = 0;
  for (int r = 0; r < self; r++) n_before += counts[r];
  u64 *before = b_fp_before_.get<u64>(2 * std::max<u64>(n_before, 1));
  u64 pos = 0;
  for (int r = 0; r < self; r++) {
    if (counts[r])  // UVA: works across devices (peer copy when enabled, staged otherwise)
      BSK_CUDA(cudaMemcpyAsync(before + 2 * pos, e[r]->b_fp_.p, counts[r] * 16, cudaMemcpyDefault, stream));
    pos += counts[r];
  }
  return rmdup_resolve_device(before, n_before, out);
}

}  // namespace bsk
