/*
 * bsk_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the BigSeqKit per-record hot path, used as the
 * parity oracle for the CUDA library and as the "port" CPU baseline in bench.py.
 * Nothing under bigseqkit_b200/ may include, link or call this.
 *
 * PARITY UNPINNED: the reference (citiususc/BigSeqKit @3ab4862) ships no tests or
 * fixtures and cannot be built here (no Go toolchain; IgnisHPC and the
 * shenwei356/bio v0.7.0, shenwei356/util v0.5.0, cespare/xxhash/v2 v2.1.2
 * modules are absent).  This file follows the reference's control flow
 * (the .go files of bigseqkit-lib/, cited per function) and restates the published
 * algorithms of the third-party leaves; the only external pins are the
 * help-text tables (bigseqkit-cli/helper.go:348-361, translate.go:42-52), the
 * public XXH64 test vectors and the NCBI genetic-code strings.
 */
#ifndef BSK_ORACLE_H
#define BSK_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Flat view of the reference option structs (the .go files of bigseqkit/: KitConfig,
 * SeqOptions, StatsOptions, RmDupOptions, TranslateOptions, LocateOptions,
 * GrepOptions, SubseqOptions).  Field names follow the Go fields. */
typedef struct {
  /* KitConfig (bigseqkit/helper.go:28-38, defaults :86-103) */
  const char *SeqType;          /* "auto" */
  int LineWidth;                /* 60 */
  int IDNCBI;                   /* 0 ; only default IDRegexp and --id-ncbi are supported */
  int AlphabetGuessSeqLength;   /* 10000 */
  /* SeqOptions (bigseqkit/seq.go:9-55) */
  int Reverse, Complement, Name, Seq, Qual, OnlyId, RemoveGaps;
  const char *GapLetters;       /* "- \t." (seq) / "- ." (stats) */
  int LowerCase, UpperCase, Dna2rna, Rna2dna, ValidateSeq, ValidateSeqLength;
  int MaxLen, MinLen, QualAsciiBase;
  double MinQual, MaxQual;
  /* StatsOptions (bigseqkit/stats.go:18-38) */
  int Tabular, All;
  const char *FqEncoding;       /* "sanger" */
  /* RmDupOptions (bigseqkit/rmdup.go:13-33) ; Grep shares ByName/BySeq/IgnoreCase/OnlyPositiveStrand */
  int ByName, BySeq, IgnoreCase, OnlyPositiveStrand;
  /* TranslateOptions (bigseqkit/translate.go:9-35) */
  int TranslTable;
  const char *Frame;            /* CSV, "1" */
  int Trim, Clean, AllowUnknownCodon, InitCodonAsM, AppendFrame;
  /* LocateOptions / GrepOptions (bigseqkit/locate.go:9-45, grep.go:13-49) */
  int n_patterns;
  const char *const *pattern_names; /* locate: name column; grep: unused */
  const char *const *patterns;
  int NonGreedy, Gtf, Bed, HideMatched, Circular, InvertMatch, Count;
  /* SubseqOptions / Grep region (bigseqkit/subseq.go:9-35) */
  const char *Region;           /* "" */
} orc_opts;

typedef struct {
  uint8_t *data;        /* every element followed by one '\n' (lib/helper.go:447) */
  size_t n;
  uint64_t *elem_off;   /* n_elem+1 offsets into data (element i = [off[i], off[i+1]-1)) */
  size_t n_elem;
  char err[512];        /* non-empty => the operator returned an error */
} orc_out;

typedef struct {
  uint64_t num, sum_len, min_len, max_len, sum_gap, q20, q30, n50, l50;
  double avg_len, q1, q2, q3, q20_pct, q30_pct;
  char type[16];        /* "DNA","RNA","Protein","Unlimit","" ... */
  /* sparse histogram, ascending length */
  uint64_t *hist_len, *hist_cnt;
  size_t n_hist;
  char err[512];
} orc_stats;

void orc_opts_default(orc_opts *o);
void orc_out_free(orc_out *o);
void orc_stats_free(orc_stats *s);

/* framing: drv/helper.go:148-178 + lib/helper.go:41-66; returns count, *starts malloc'd (n+1 entries, last = n) */
size_t orc_frame(const uint8_t *data, size_t n, uint64_t **starts);

uint64_t orc_xxh64(const uint8_t *p, size_t n, uint64_t seed);

int orc_seq(const uint8_t *data, size_t n, const orc_opts *o, orc_out *out);
int orc_stats_run(const uint8_t *data, size_t n, const orc_opts *o, orc_stats *out);
/* merge (sum semantics) and finalise several partial results */
int orc_stats_merge(orc_stats *dst, const orc_stats *src);
void orc_stats_finalise(orc_stats *s, int all);
/* renders StatsString (drv/stats.go:168-288); returns malloc'd NUL-terminated string */
char *orc_stats_render(const orc_stats *s, const char *file, const char *format, int tabular, int all);
int orc_rmdup(const uint8_t *data, size_t n, const orc_opts *o, orc_out *out, uint64_t *n_removed);
/* -d / -D side outputs of RmDupCheck (lib/rmdup.go:180-239): removed records; "count\tid1, id2, ..." rows */
int orc_rmdup_dups(const uint8_t *data, size_t n, const orc_opts *o, orc_out *dup_seqs, orc_out *dup_num);
/* keys only (RmDupPrepare, lib/rmdup.go:43-90): int64 keys per record */
int orc_rmdup_keys(const uint8_t *data, size_t n, const orc_opts *o, int64_t **keys, size_t *n_keys);
int orc_translate(const uint8_t *data, size_t n, const orc_opts *o, orc_out *out);
int orc_locate(const uint8_t *data, size_t n, const orc_opts *o, orc_out *out);
int orc_grep(const uint8_t *data, size_t n, const orc_opts *o, orc_out *out);
int orc_subseq(const uint8_t *data, size_t n, const orc_opts *o, orc_out *out);
/* Fq2Fa.Call (lib/fq2fa.go:36-61) */
int orc_fq2fa(const uint8_t *data, size_t n, const orc_opts *o, orc_out *out);
/* raw-element operators (lib/duplicate.go:24-30, lib/range.go:26-44, bigseqkit/head.go:33-44) */
int orc_duplicate(const uint8_t *data, size_t n, int64_t times, orc_out *out);
int orc_range(const uint8_t *data, size_t n, int64_t start, int64_t end, int64_t index_base, orc_out *out);

/* leaf helpers exposed for known-answer tests */
size_t orc_subseq_range(size_t len, int start, int end, size_t *s0); /* returns length, *s0 = 0-based start */
int orc_translate_codon(int table, const char *codon);                /* aa or -1 unknown */
size_t orc_wrap_len(size_t l, int width);

/* multi-threaded CPU baseline: record-aligned shards, one thread each; op in
 * {"seq","stats","rmdup"}; returns records processed, fills checksum of output sizes */
int orc_run_mt(const char *op, const uint8_t *data, size_t n, const orc_opts *o, int threads,
               uint64_t *n_records, uint64_t *out_bytes);
/* same sharded run, output kept: op in {"seq","translate","locate","grep","subseq","fq2fa","rmdup"} fills *out with the
 * elements of all shards in input order (locate: header row from shard 0 only, lib/locate.go:198-204; rmdup: first
 * occurrence in input order over the WHOLE input); op "stats" fills *st with the merged, finalised totals. */
int orc_run_mt_out(const char *op, const uint8_t *data, size_t n, const orc_opts *o, int threads, orc_out *out,
                   orc_stats *st, uint64_t *n_records);

#ifdef __cplusplus
}
#endif
#endif
