/*
 * bsk.h -- C ABI of libbsk.so, the B200-native engine behind BigSeqKit's per-record
 * operators (seq / stats / subseq / grep / locate / rmdup / translate, and fq2fa as the
 * first of the operators either side of that path).
 *
 * Every entry point replaces one piece of the reference's executor-plugin surface
 * (citations are relative to the reference tree, citiususc/BigSeqKit @3ab4862):
 *
 *   reference                                                   | here
 *   ------------------------------------------------------------+---------------------------
 *   exported factory  New<Name>() any                           | bsk_create("<Name>", opts)
 *     bigseqkit-lib/seq.go:17-19, stats.go:16,119,              |
 *     rmdup.go:23,92, translate.go:21, locate.go:19,            |
 *     grep.go:24, subseq.go:22, fq2fa.go:15                     |
 *   Before(ctx): opts := StringToOptions(ctx.Vars()["opts"])    | bsk_create parses the same JSON
 *     bigseqkit-lib/seq.go:28-79, bigseqkit/helper.go:47-66     | and returns the same error text
 *   Call(it IReadIterator[string], ctx) ([]string, error)       | bsk_run_buffer / bsk_run_device
 *     bigseqkit-lib/seq.go:81 (IMapPartitions)                  |   (one call == one partition)
 *   Call(pid int64, it, ctx) (IMapPartitionsWithIndex)          | same, partition_id argument
 *     bigseqkit-lib/locate.go:195, grep.go:544                  |
 *   StatsReduce.Call(v1, v2 map[int64]int64)                    | bsk_stats_merge / bsk_stats_add
 *     bigseqkit-lib/stats.go:128-137                            |
 *   Stats finalise + StatsString  bigseqkit/stats.go:75-288     | bsk_stats_result / bsk_stats_render
 *   RmDupPrepare key = int64(xxhash.Sum64(subject))             | bsk_rmdup_keys* (+ "RmDup" fused op)
 *     bigseqkit-lib/rmdup.go:67-86                              |
 *   Stats Reduce over all partitions  bigseqkit/stats.go:91     | bsk_stats_allreduce (NCCL) / bsk_reduce
 *   RmDup GroupByKey + RmDupCheck     bigseqkit/rmdup.go:97,    | bsk_rmdup_sharded (NCCL) / bsk_rmdup_union
 *     bigseqkit-lib/rmdup.go:118-242                            |
 *   FileStore token ring  bigseqkit-lib/helper.go:399-431       | bsk_output_offsets
 *   After(ctx) / plugin unload                                  | bsk_destroy
 *   error return of Before/Call                                 | negative status + bsk_last_error
 *
 * Conventions: plain C types only; no C++/torch types; no exceptions cross the
 * boundary; a ctx is bound to one CUDA device and is used by one thread at a
 * time (the reference calls Call() concurrently per executor thread -> use one
 * ctx per thread).  Input = the bytes of one partition: whole FASTA/FASTQ
 * records, i.e. exactly what worker.PlainFile + ReadFixer hand to Call()
 * (bigseqkit/helper.go:148-178, bigseqkit-lib/helper.go:41-66) concatenated
 * with their '\n' separators.  Output = the returned []string written the way
 * FileStore writes it: every element followed by one '\n'
 * (bigseqkit-lib/helper.go:441-451), plus element offsets so that a binding can
 * slice it back into strings without copying.
 */
#ifndef BSK_H
#define BSK_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSK_OK 0
#define BSK_ERR_ARG (-1)         /* option / flag validation failed (reference: Before() error) */
#define BSK_ERR_DATA (-2)        /* malformed input (reference: Call() error) */
#define BSK_ERR_CUDA (-3)        /* CUDA runtime failure or no usable device */
#define BSK_ERR_UNSUPPORTED (-4) /* flag combination outside the accelerated path */
#define BSK_ERR_STATE (-5)       /* call sequence error, or an internal error that is neither CUDA nor data */
#define BSK_ERR_NOMEM (-6)       /* out of host memory */

typedef struct bsk_ctx bsk_ctx;

/* Result of one Call().  For bsk_run_buffer the pointers are pinned host memory,
 * for bsk_run_device they are device memory; both stay owned by the ctx and are
 * valid until the next run/reset/destroy on the same ctx. */
typedef struct {
  uint8_t *data;      /* n bytes: element, '\n', element, '\n', ... */
  size_t n;
  uint64_t *elem_off; /* n_elem + 1 offsets into data (element i = [off[i], off[i+1]-1)); NULL unless requested */
  size_t n_elem;
  uint64_t n_records; /* input records seen by this call */
} bsk_out;

/* Stats result: bigseqkit/stats.go:75-166 (finalised row) + the raw reduce value
 * of bigseqkit-lib/stats.go:48-117 as a sparse ascending length histogram. */
typedef struct {
  uint64_t num, sum_len, min_len, max_len, sum_gap, q20, q30, n50, l50;
  double avg_len, q1, q2, q3, q20_pct, q30_pct;
  char type[16];            /* "DNA" "RNA" "Protein" "Unlimit" "" */
  const uint64_t *hist_len; /* owned by the ctx */
  const uint64_t *hist_cnt;
  size_t n_hist;
} bsk_stats;

/* Per-call device timings (CUDA events on the ctx stream), for bench/profiling. */
typedef struct {
  float index_ms;           /* delimiter scan -> record index -> parse (+ squeeze, alphabet guess) */
  float op_ms;              /* operator kernels after the index, scans included */
  float main_ms;            /* the operator's dominant kernel alone (record emitter / hash / matcher) */
  float total_ms;           /* whole call on the ctx stream; host<->device copies included for bsk_run_buffer */
  uint64_t main_launches;   /* launches of the dominant kernel summed in main_ms */
  uint64_t kernel_launches; /* launches of libbsk's own kernels in the last call */
  uint64_t in_bytes, out_bytes;
  uint64_t fused_blocks;    /* blocks handled by the single-pass tile kernels (short-record path) */
} bsk_timings;

int bsk_version(void);
int bsk_device_count(void);

/* op: "SeqTransform" | "Stats" | "RmDup" | "RmDupPrepare" | "Translate" | "Locate" | "Grep" | "SubseqTransform".
 * opts_json: the reference's option struct as JSON, e.g.
 *   {"Config":{"SeqType":"auto","LineWidth":60,...},"Reverse":true,"Complement":true}
 * missing fields take the reference defaults (bigseqkit/helper.go:86-103, seq.go:32-55, ...).
 * device < 0 selects the current device.  On failure *out is NULL and
 * bsk_create_error() holds the message. */
int bsk_create(const char *op, const char *opts_json, int device, bsk_ctx **out);
const char *bsk_create_error(void);
void bsk_destroy(bsk_ctx *ctx);
const char *bsk_last_error(const bsk_ctx *ctx);
/* want_elem_off != 0: fill bsk_out.elem_off (default on). */
int bsk_set_elem_offsets(bsk_ctx *ctx, int want_elem_off);
/* on != 0: the partitions of the following calls are ONE dataframe (bigseqkit-cli/helper.go:131-138 Unions all input
 * files before the operator runs; bigseqkit/rmdup.go:97 groups keys over all partitions): the rmdup key table and
 * history and the RangePrepare record index (bigseqkit-lib/range.go:26-31, MapWithIndex) keep running from call to
 * call, in call order, until bsk_reset.  Default off: every call is a partition of its own. */
int bsk_set_union(bsk_ctx *ctx, int on);
/* forget accumulated state (stats totals, rmdup keys) */
int bsk_reset(bsk_ctx *ctx);

/* One Call() on a partition held in HOST memory (pageable or pinned).  The input is
 * staged to HBM in record-aligned blocks with cudaMemcpyAsync overlapping the
 * kernels, results come back into the ctx's pinned arena. */
int bsk_run_buffer(bsk_ctx *ctx, const uint8_t *in, size_t n, int64_t partition_id, bsk_out *out);
/* Same Call() on a partition already resident in HBM (d_in: device pointer,
 * n < 4 GiB - 64); out->data / out->elem_off are device pointers. */
int bsk_run_device(bsk_ctx *ctx, const void *d_in, size_t n, int64_t partition_id, bsk_out *out);
/* File-level entry points -- what a driver that reads files itself (the reference's ReadFASTA / ReadFASTQ + StoreFASTX,
 * bigseqkit/helper.go:148-195, bigseqkit-lib/helper.go:378-460) calls instead of handing Go strings around.
 *
 * bsk_shard_bounds: cut the file into n_shards contiguous record-aligned byte ranges (every cut moves forward to the
 *   next record start: FASTA a line starting '>', FASTQ a line starting '@' that does not follow a bare "+" line);
 *   bounds receives n_shards + 1 offsets.  This is the job worker.PlainFile(path, delim) does for the reference.
 * bsk_run_file: one Call() on the byte range [off, off + len) of `path` (len == 0: to the end of the file), which must
 *   start on a record.  The range is read into pinned memory, run through the same three-stream pipeline as
 *   bsk_run_buffer, and the result (every element followed by '\n') is written to out_path at byte offset
 *   out_off (the file is created if needed, never truncated), so that ranks can write one merged file at the
 *   offsets an all-gather of their sizes gave them.  out_path == NULL: nothing is written (stats).
 *   *out_bytes / *n_records / *n_elem receive the totals when not NULL. */
int bsk_shard_bounds(const char *path, int n_shards, uint64_t *bounds);
int bsk_run_file(bsk_ctx *ctx, const char *path, uint64_t off, uint64_t len, int64_t partition_id, const char *out_path,
                 uint64_t out_off, uint64_t *out_bytes, uint64_t *n_records, uint64_t *n_elem);
/* cudaStream_t the ctx launches on (as void*), for callers that time with CUDA events. */
void *bsk_stream(bsk_ctx *ctx);
int bsk_get_timings(const bsk_ctx *ctx, bsk_timings *t);

/* ---- stats -------------------------------------------------------------- */
/* totals accumulated over every run_* since create/reset, sum semantics */
int bsk_stats_result(bsk_ctx *ctx, bsk_stats *out);
/* StatsReduce: dst += src */
int bsk_stats_merge(bsk_ctx *dst, const bsk_ctx *src);
/* add a partial result that travelled as plain arrays (e.g. after an NCCL all-gather) */
int bsk_stats_add(bsk_ctx *ctx, const uint64_t *hist_len, const uint64_t *hist_cnt, size_t n_hist,
                  uint64_t q20, uint64_t q30, uint64_t sum_gap, const char *type);
/* dense device histogram of the accumulated lengths < nbins (uint64 counts) for an
 * NCCL all-reduce; lengths >= nbins are reported through *n_overflow */
int bsk_stats_dense_device(bsk_ctx *ctx, void *d_hist_u64, size_t nbins, uint64_t *n_overflow);
/* StatsString (bigseqkit/stats.go:168-288): returns the number of bytes needed
 * (excluding NUL); writes at most cap bytes */
long bsk_stats_render(bsk_ctx *ctx, const char *file, const char *format, char *buf, size_t cap);

/* ---- rmdup -------------------------------------------------------------- */
/* keys of the last "RmDup"/"RmDupPrepare" call, int64(xxhash.Sum64(subject)), one per input record */
int bsk_rmdup_keys(bsk_ctx *ctx, const int64_t **keys, size_t *n);
/* number of records dropped as duplicates by the last "RmDup" call */
uint64_t bsk_rmdup_removed(const bsk_ctx *ctx);
/* rmdup -d / -D (options "DupSeqsFile" / "DupNumFile" non-empty; RmDupCheck.Call bigseqkit-lib/rmdup.go:180-239,
 * written by After :245-275).  Text accumulated over the bsk_run_* calls since the last partition start, owned by
 * the ctx until the next call on it:
 *   dup_seqs  every removed record as Record.Format(LineWidth), input order;
 *   dup_num   "count\tid1, id2, ...\n" per subject with more than one member, rows ordered by the group's first
 *             member (the reference walks a Go map there), ids in input order.
 * The caller writes the files; the flag meaning is implemented, not the directory swap of rmdup.go:246-267. */
int bsk_rmdup_dup_seqs(bsk_ctx *ctx, const char **data, size_t *n);
int bsk_rmdup_dup_num(bsk_ctx *ctx, const char **data, size_t *n);
/* Multi-GPU rmdup (the GroupByKey exchange of bigseqkit/rmdup.go:97 as one all-gather):
 *  1. bsk_rmdup_prepare_device: index + hash the local shard; d_fp receives n_records
 *     16-byte fingerprints {xxh64 seed 0, xxh64 seed 0x9E3779B97F4A7C15 ^ len};
 *  2. the caller all-gathers the fingerprints of all ranks (rank order == input order);
 *  3. bsk_rmdup_resolve_device: d_all = fingerprints of every record before and
 *     including this shard (n_before + n_local entries); marks local records that
 *     have an earlier equal fingerprint and emits the survivors. */
int bsk_rmdup_prepare_device(bsk_ctx *ctx, const void *d_in, size_t n, void *d_fp, size_t fp_cap, uint64_t *n_records);
int bsk_rmdup_resolve_device(bsk_ctx *ctx, const void *d_all_fp, uint64_t n_before, bsk_out *out);

/* ---- exchange steps between partitions ------------------------------------- */
/* The path has exactly two: StatsReduce (bigseqkit/stats.go:91, bigseqkit-lib/stats.go:128-137; sum semantics) and
 * rmdup's GroupByKey + RmDupCheck (bigseqkit/rmdup.go:97, bigseqkit-lib/rmdup.go:118-242; the first occurrence in
 * global input order survives).  Plus the ordered merged output file, for which the reference passes an MPI token
 * around its executors (bigseqkit-lib/helper.go:399-431).
 *
 * One process per GPU (the reference: one executor per MPI rank): rank 0 calls bsk_comm_unique_id, the driver ships
 * the 128 bytes to every rank (MPI / IgnisHPC variable / torch.distributed / a file), every rank calls
 * bsk_comm_init on its ctx (rank order == input order).  The collectives below run on the ctx stream over NCCL
 * (NVLink / NVSwitch); NCCL is loaded on first use (libnccl.so.2, or $BSK_NCCL_LIB). */
#define BSK_COMM_ID_BYTES 128
int bsk_comm_unique_id(uint8_t *id /* BSK_COMM_ID_BYTES */);
const char *bsk_comm_error(void); /* message of the last failed bsk_comm_unique_id on this thread */
int bsk_comm_init(bsk_ctx *ctx, const uint8_t *id, int n_ranks, int rank);
int bsk_comm_rank(const bsk_ctx *ctx, int *rank, int *n_ranks);
int bsk_comm_free(bsk_ctx *ctx);
/* byte offset of this rank's output in the merged file and the total size: one all-gather of the local sizes */
int bsk_output_offsets(bsk_ctx *ctx, uint64_t n_local, uint64_t *offset, uint64_t *total);
/* Stats: after the call every rank's totals (bsk_stats_result / bsk_stats_render) are the global ones.  One
 * all-reduce of a dense 65536-bin length histogram + the Q20/Q30/gap sums, one small all-gather for the record
 * counts / type column / lengths beyond the dense range. */
int bsk_stats_allreduce(bsk_ctx *ctx);
/* RmDup over the shards of all ranks: hash the local shard (d_in: device pointer, whole records), all-gather the
 * 16-byte fingerprints, drop every local record with an equal fingerprint earlier in global order, emit the
 * survivors (out: device pointers).  Inside a shard equal keys are confirmed by comparing the subject bytes, as
 * RmDupCheck does; across shards the 128-bit fingerprint {XXH64 seed 0, XXH64 seed 0x9E3779B97F4A7C15 ^ length}
 * decides (the bytes live on another GPU).  Every rank must call it; a rank whose shard fails makes all ranks
 * return an error. */
int bsk_rmdup_sharded(bsk_ctx *ctx, const void *d_in, size_t n, bsk_out *out);
/* Several partitions inside ONE process (the reference runs Call() once per partition on executor threads): no
 * NCCL, plain device copies; ctx order == input order; the ctxs may sit on different devices.
 *   bsk_reduce       StatsReduce folded over n Stats ctxs: every ctx ends with the sum of all;
 *   bsk_rmdup_union  rmdup over n shards (d_in[i], n_bytes[i] on ctxs[i]'s device), outs[i] = survivors of shard i. */
int bsk_reduce(bsk_ctx **ctxs, int n);
int bsk_rmdup_union(bsk_ctx **ctxs, int n, const void *const *d_in, const size_t *n_bytes, bsk_out *outs);
/* copy a host partition (< 4 GiB - 1 MiB) into a device buffer owned by the ctx; *d_ptr stays valid until the next
 * bsk_stage_device / bsk_run_buffer on the ctx.  For bindings that feed bsk_run_device / bsk_rmdup_sharded without
 * touching the CUDA runtime themselves (the Go shim). */
int bsk_stage_device(bsk_ctx *ctx, const uint8_t *in, size_t n, void **d_ptr);
/* copy n bytes of a device result (bsk_run_device / bsk_rmdup_* outputs) to host memory, ordered after the ctx stream */
int bsk_memcpy_d2h(bsk_ctx *ctx, void *h_dst, const void *d_src, size_t n);

/* ---- grep --------------------------------------------------------------- */
/* matched-record count of the last "Grep" call (also delivered as the element when Count is set) */
uint64_t bsk_grep_count(const bsk_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
