// api.cu -- extern "C" boundary of libbsk.so (declarations: include/bsk.h).
#include <cstring>
#include <new>
#include <string>

#include "bsk.h"
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "engine.h"

struct bsk_ctx {
  bsk::Engine *eng;
};

static thread_local std::string g_create_err;

// exceptions never cross the boundary; each class keeps its own status code
#define BSK_GUARD(ctx, body)                                      \
  try {                                                           \
    body                                                          \
  } catch (const bsk::CudaError &e) {                             \
    (ctx)->eng->err = e.what();                                   \
    return BSK_ERR_CUDA;                                          \
  } catch (const std::bad_alloc &) {                              \
    (ctx)->eng->err = "out of host memory";                       \
    return BSK_ERR_NOMEM;                                         \
  } catch (const std::invalid_argument &e) {                      \
    (ctx)->eng->err = e.what();                                   \
    return BSK_ERR_ARG;                                           \
  } catch (const std::exception &e) {                             \
    (ctx)->eng->err = std::string("internal error: ") + e.what(); \
    return BSK_ERR_STATE;                                         \
  }

extern "C" {

int bsk_version(void) { return 100; }

int bsk_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char *bsk_create_error(void) { return g_create_err.c_str(); }

int bsk_create(const char *op, const char *opts_json, int device, bsk_ctx **out) {
  if (!out) return BSK_ERR_ARG;
  *out = nullptr;
  g_create_err.clear();
  const bsk::Op o = bsk::op_from_name(op);
  if (o == bsk::OP_INVALID) {
    g_create_err = std::string("unknown operator: ") + (op ? op : "(null)");
    return BSK_ERR_ARG;
  }
  bsk::Opts opts;
  int code = BSK_OK;
  if (!bsk::parse_and_validate(o, opts_json, opts, g_create_err, code)) return code;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    g_create_err = "no CUDA device available: libbsk has no CPU fallback";
    return BSK_ERR_CUDA;
  }
  if (device >= ndev) {
    g_create_err = "device index out of range";
    return BSK_ERR_ARG;
  }
  try {
    bsk_ctx *c = new bsk_ctx;
    c->eng = new bsk::Engine(o, opts, device);
    *out = c;
  } catch (const std::exception &e) {
    g_create_err = e.what();
    return BSK_ERR_CUDA;
  }
  return BSK_OK;
}

void bsk_destroy(bsk_ctx *ctx) {
  if (!ctx) return;
  delete ctx->eng;
  delete ctx;
}

const char *bsk_last_error(const bsk_ctx *ctx) { return ctx ? ctx->eng->err.c_str() : "null ctx"; }

int bsk_set_elem_offsets(bsk_ctx *ctx, int want) {
  if (!ctx) return BSK_ERR_ARG;
  ctx->eng->want_elem_off = want ? 1 : 0;
  return BSK_OK;
}

int bsk_set_union(bsk_ctx *ctx, int on) {
  if (!ctx) return BSK_ERR_ARG;
  ctx->eng->set_union(on != 0);
  return BSK_OK;
}

int bsk_reset(bsk_ctx *ctx) {
  if (!ctx) return BSK_ERR_ARG;
  BSK_GUARD(ctx, return ctx->eng->reset();)
}

int bsk_run_buffer(bsk_ctx *ctx, const uint8_t *in, size_t n, int64_t partition_id, bsk_out *out) {
  if (!ctx || !out || (!in && n)) return BSK_ERR_ARG;
  ctx->eng->err.clear();
  BSK_GUARD(ctx, return ctx->eng->run_buffer(in, n, partition_id, out);)
}

int bsk_run_device(bsk_ctx *ctx, const void *d_in, size_t n, int64_t partition_id, bsk_out *out) {
  if (!ctx || !out || (!d_in && n)) return BSK_ERR_ARG;
  ctx->eng->err.clear();
  BSK_GUARD(ctx, return ctx->eng->run_device(d_in, n, partition_id, out);)
}

// ---- file-level entry points (host code: POSIX I/O around bsk_run_buffer)
static bool record_start_at(const std::vector<unsigned char> &w, size_t i, bool fq) {
  // w holds bytes [base, base + w.size()); i >= 3 is an index into it
  const unsigned char marker = fq ? '@' : '>';
  if (w[i - 1] != '\n' || w[i] != marker) return false;
  if (fq && w[i - 3] == '\n' && w[i - 2] == '+') return false;
  return true;
}

int bsk_shard_bounds(const char *path, int n_shards, uint64_t *bounds) {
  if (!path || n_shards < 1 || !bounds) return BSK_ERR_ARG;
  FILE *f = fopen(path, "rb");
  if (!f) return BSK_ERR_ARG;
  fseeko(f, 0, SEEK_END);
  const uint64_t n = (uint64_t)ftello(f);
  int first = 0;
  if (n) { fseeko(f, 0, SEEK_SET); first = fgetc(f); }
  const bool fq = first == '@';
  bounds[0] = 0;
  std::vector<unsigned char> w;
  for (int r = 1; r < n_shards; r++) {
    uint64_t pos = n / (uint64_t)n_shards * (uint64_t)r;
    if (pos < bounds[r - 1]) pos = bounds[r - 1];
    uint64_t cut = n;
    // scan forward in 1 MiB windows (3 bytes of look-behind) for the next record start
    while (pos < n) {
      const uint64_t base = pos >= 3 ? pos - 3 : 0;
      const size_t want = (size_t)std::min<uint64_t>(n - base, (1u << 20) + 3);
      w.resize(want);
      fseeko(f, (off_t)base, SEEK_SET);
      if (fread(w.data(), 1, want, f) != want) { fclose(f); return BSK_ERR_DATA; }
      bool found = false;
      for (size_t i = (size_t)(pos - base); i < want; i++) {
        if (base + i == 0) { cut = 0; found = true; break; }
        if (i >= 3 ? record_start_at(w, i, fq) : (base + i >= 1 && w[i - 1] == '\n' && w[i] == (fq ? '@' : '>'))) {
          cut = base + i;
          found = true;
          break;
        }
      }
      if (found) break;
      pos = base + want;
    }
    bounds[r] = cut;
  }
  bounds[n_shards] = n;
  fclose(f);
  return BSK_OK;
}

int bsk_run_file(bsk_ctx *ctx, const char *path, uint64_t off, uint64_t len, int64_t partition_id, const char *out_path,
                 uint64_t out_off, uint64_t *out_bytes, uint64_t *n_records, uint64_t *n_elem) {
  if (!ctx || !path) return BSK_ERR_ARG;
  ctx->eng->err.clear();
  const int fd = open(path, O_RDONLY);
  if (fd < 0) { ctx->eng->err = std::string("cannot open ") + path; return BSK_ERR_ARG; }
  struct stat sb;
  if (fstat(fd, &sb) != 0) { close(fd); ctx->eng->err = std::string("cannot stat ") + path; return BSK_ERR_ARG; }
  const uint64_t fsz = (uint64_t)sb.st_size;
  if (off > fsz) off = fsz;
  if (len == 0 || off + len > fsz) len = fsz - off;
  int g = -1;
  if (out_path) {
    g = open(out_path, O_RDWR | O_CREAT, 0644);
    if (g < 0) { close(fd); ctx->eng->err = std::string("cannot open ") + out_path; return BSK_ERR_ARG; }
  }
  struct Closer {
    int a, b;
    ~Closer() { close(a); if (b >= 0) close(b); }
  } closer{fd, g};
  BSK_GUARD(ctx, return ctx->eng->run_stream(fd, off, len, partition_id, g, out_off, out_bytes, n_records, n_elem);)
}

void *bsk_stream(bsk_ctx *ctx) { return ctx ? (void *)ctx->eng->stream : nullptr; }

int bsk_get_timings(const bsk_ctx *ctx, bsk_timings *t) {
  if (!ctx || !t) return BSK_ERR_ARG;
  *t = ctx->eng->timings;
  return BSK_OK;
}

int bsk_stats_result(bsk_ctx *ctx, bsk_stats *out) {
  if (!ctx || !out) return BSK_ERR_ARG;
  BSK_GUARD(ctx, return ctx->eng->stats_result(out);)
}

int bsk_stats_merge(bsk_ctx *dst, const bsk_ctx *src) {
  if (!dst || !src) return BSK_ERR_ARG;
  BSK_GUARD(dst, return dst->eng->stats_merge_from(*src->eng);)
}

int bsk_stats_add(bsk_ctx *ctx, const uint64_t *hist_len, const uint64_t *hist_cnt, size_t n_hist, uint64_t q20,
                  uint64_t q30, uint64_t sum_gap, const char *type) {
  if (!ctx) return BSK_ERR_ARG;
  BSK_GUARD(ctx, return ctx->eng->stats_add(hist_len, hist_cnt, n_hist, q20, q30, sum_gap, type);)
}

int bsk_stats_dense_device(bsk_ctx *ctx, void *d_hist_u64, size_t nbins, uint64_t *n_overflow) {
  if (!ctx || !d_hist_u64) return BSK_ERR_ARG;
  BSK_GUARD(ctx, return ctx->eng->stats_dense_device(d_hist_u64, nbins, n_overflow);)
}

long bsk_stats_render(bsk_ctx *ctx, const char *file, const char *format, char *buf, size_t cap) {
  if (!ctx) return -1;
  try {
    return ctx->eng->stats_render(file ? file : "", format ? format : "", buf, cap);
  } catch (const std::exception &e) {
    ctx->eng->err = e.what();
    return -1;
  }
}

int bsk_rmdup_keys(bsk_ctx *ctx, const int64_t **keys, size_t *n) {
  if (!ctx || !keys || !n) return BSK_ERR_ARG;
  BSK_GUARD(ctx, return ctx->eng->rmdup_keys(keys, n);)
}

int bsk_rmdup_dup_seqs(bsk_ctx *ctx, const char **data, size_t *n) {
  if (!ctx || !data || !n) return BSK_ERR_ARG;
  BSK_GUARD(ctx, return ctx->eng->rmdup_dup_seqs(data, n);)
}

int bsk_rmdup_dup_num(bsk_ctx *ctx, const char **data, size_t *n) {
  if (!ctx || !data || !n) return BSK_ERR_ARG;
  BSK_GUARD(ctx, return ctx->eng->rmdup_dup_num(data, n);)
}

uint64_t bsk_rmdup_removed(const bsk_ctx *ctx) { return ctx ? ctx->eng->rmdup_removed : 0; }

int bsk_rmdup_prepare_device(bsk_ctx *ctx, const void *d_in, size_t n, void *d_fp, size_t fp_cap, uint64_t *n_records) {
  if (!ctx) return BSK_ERR_ARG;
  ctx->eng->err.clear();
  BSK_GUARD(ctx, return ctx->eng->rmdup_prepare_device(d_in, n, d_fp, fp_cap, n_records);)
}

int bsk_rmdup_resolve_device(bsk_ctx *ctx, const void *d_all_fp, uint64_t n_before, bsk_out *out) {
  if (!ctx || !out) return BSK_ERR_ARG;
  ctx->eng->err.clear();
  BSK_GUARD(ctx, return ctx->eng->rmdup_resolve_device(d_all_fp, n_before, out);)
}

uint64_t bsk_grep_count(const bsk_ctx *ctx) { return ctx ? ctx->eng->grep_count : 0; }

int bsk_memcpy_d2h(bsk_ctx *ctx, void *h_dst, const void *d_src, size_t n) {
  if (!ctx || (n && (!h_dst || !d_src))) return BSK_ERR_ARG;
  BSK_GUARD(ctx, {
    if (n) BSK_CUDA(cudaMemcpyAsync(h_dst, d_src, n, cudaMemcpyDeviceToHost, ctx->eng->stream));
    BSK_CUDA(cudaStreamSynchronize(ctx->eng->stream));
    return BSK_OK;
  })
}

int bsk_stage_device(bsk_ctx *ctx, const uint8_t *in, size_t n, void **d_ptr) {
  if (!ctx || !d_ptr || (!in && n)) return BSK_ERR_ARG;
  ctx->eng->err.clear();
  BSK_GUARD(ctx, return ctx->eng->stage_device(in, n, d_ptr);)
}

// ---- exchange steps
static thread_local std::string g_comm_err;
const char *bsk_comm_error(void) { return g_comm_err.c_str(); }

int bsk_comm_unique_id(uint8_t *id) {
  if (!id) return BSK_ERR_ARG;
  g_comm_err.clear();
  try {
    return bsk::comm_unique_id(id, g_comm_err);
  } catch (const std::exception &e) {
    g_comm_err = e.what();
    return BSK_ERR_CUDA;
  }
}

int bsk_comm_init(bsk_ctx *ctx, const uint8_t *id, int n_ranks, int rank) {
  if (!ctx || !id) return BSK_ERR_ARG;
  ctx->eng->err.clear();
  BSK_GUARD(ctx, return ctx->eng->comm_init(id, n_ranks, rank);)
}

int bsk_comm_free(bsk_ctx *ctx) {
  if (!ctx) return BSK_ERR_ARG;
  BSK_GUARD(ctx, { ctx->eng->comm_free(); return BSK_OK; })
}

int bsk_comm_rank(const bsk_ctx *ctx, int *rank, int *n_ranks) {
  if (!ctx) return BSK_ERR_ARG;
  return ctx->eng->comm_rank(rank, n_ranks);
}

int bsk_output_offsets(bsk_ctx *ctx, uint64_t n_local, uint64_t *offset, uint64_t *total) {
  if (!ctx) return BSK_ERR_ARG;
  ctx->eng->err.clear();
  BSK_GUARD(ctx, return ctx->eng->output_offsets(n_local, offset, total);)
}

int bsk_stats_allreduce(bsk_ctx *ctx) {
  if (!ctx) return BSK_ERR_ARG;
  ctx->eng->err.clear();
  BSK_GUARD(ctx, return ctx->eng->stats_allreduce();)
}

int bsk_rmdup_sharded(bsk_ctx *ctx, const void *d_in, size_t n, bsk_out *out) {
  if (!ctx || !out || (!d_in && n)) return BSK_ERR_ARG;
  ctx->eng->err.clear();
  BSK_GUARD(ctx, return ctx->eng->rmdup_sharded(d_in, n, out);)
}

int bsk_reduce(bsk_ctx **ctxs, int n) {
  if (!ctxs || n < 1) return BSK_ERR_ARG;
  std::vector<bsk::Engine *> e;
  for (int i = 0; i < n; i++) {
    if (!ctxs[i]) return BSK_ERR_ARG;
    e.push_back(ctxs[i]->eng);
  }
  BSK_GUARD(ctxs[0], {
    std::string err;
    const int rc = bsk::stats_reduce_local(e.data(), n, err);
    if (rc != BSK_OK) ctxs[0]->eng->err = err;
    return rc;
  })
}

int bsk_rmdup_union(bsk_ctx **ctxs, int n, const void *const *d_in, const size_t *n_bytes, bsk_out *outs) {
  if (!ctxs || n < 1 || !d_in || !n_bytes || !outs) return BSK_ERR_ARG;
  std::vector<bsk::Engine *> e;
  for (int i = 0; i < n; i++) {
    if (!ctxs[i]) return BSK_ERR_ARG;
    e.push_back(ctxs[i]->eng);
  }
  BSK_GUARD(ctxs[0], {
    std::string err;
    const int rc = bsk::rmdup_union_local(e.data(), n, d_in, n_bytes, outs, err);
    if (rc != BSK_OK) ctxs[0]->eng->err = err;
    return rc;
  })
}

}  // extern "C"
