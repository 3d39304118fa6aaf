/*
 * bsk_oracle.c -- TEST INFRASTRUCTURE ONLY (see bsk_oracle.h).  PARITY UNPINNED.
 *
 * Record-at-a-time CPU restatement of the reference hot path.  Pass structure
 * and data structures follow the Go code (parse -> per-record op -> format;
 * hash map for stats/rmdup; per-pattern substring loops + per-pattern
 * reverse-complement for locate) so that it can double as the CPU baseline.
 * Citations are relative to /root/reference.  Leaves that live in un-vendored
 * modules (shenwei356/bio v0.7.0, shenwei356/util v0.5.0, cespare/xxhash v2.1.2)
 * are restated from their published behaviour and marked UNVERIFIED.
 */
#define _GNU_SOURCE
#include "bsk_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ buffers */
typedef struct { uint8_t *p; size_t n, cap; } buf_t;
static void buf_reserve(buf_t *b, size_t extra) {
  if (b->n + extra <= b->cap) return;
  size_t c = b->cap ? b->cap * 2 : 256;
  while (c < b->n + extra) c *= 2;
  b->p = (uint8_t *)realloc(b->p, c);
  b->cap = c;
}
static void buf_add(buf_t *b, const void *s, size_t n) {
  buf_reserve(b, n + 1);
  if (n) memcpy(b->p + b->n, s, n);
  b->n += n;
}
static void buf_addc(buf_t *b, uint8_t c) { buf_reserve(b, 1); b->p[b->n++] = c; }
static void buf_adds(buf_t *b, const char *s) { buf_add(b, s, strlen(s)); }
static void buf_printf(buf_t *b, const char *fmt, ...) {
  char tmp[256];
  va_list ap;
  va_start(ap, fmt);
  int k = vsnprintf(tmp, sizeof tmp, fmt, ap);
  va_end(ap);
  buf_add(b, tmp, (size_t)k);
}

typedef struct { buf_t data; uint64_t *off; size_t n, cap; } sink_t;
static void sink_elem(sink_t *s, const uint8_t *p, size_t n) {
  if (s->n + 2 > s->cap) {
    s->cap = s->cap ? s->cap * 2 : 1024;
    s->off = (uint64_t *)realloc(s->off, s->cap * sizeof(uint64_t));
  }
  s->off[s->n++] = s->data.n;
  buf_add(&s->data, p, n);
  buf_addc(&s->data, '\n'); /* FileStore: element + "\n" (lib/helper.go:447) */
  s->off[s->n] = s->data.n;
}
static void sink_to_out(sink_t *s, orc_out *o) {
  if (!s->off) { s->off = (uint64_t *)calloc(1, sizeof(uint64_t)); }
  s->off[s->n] = s->data.n;
  o->data = s->data.p; o->n = s->data.n; o->elem_off = s->off; o->n_elem = s->n;
}
void orc_out_free(orc_out *o) { free(o->data); free(o->elem_off); memset(o, 0, sizeof *o); }
void orc_stats_free(orc_stats *s) { free(s->hist_len); free(s->hist_cnt); s->hist_len = s->hist_cnt = NULL; s->n_hist = 0; }

void orc_opts_default(orc_opts *o) {
  memset(o, 0, sizeof *o);
  o->SeqType = "auto"; o->LineWidth = 60; o->AlphabetGuessSeqLength = 10000; /* bigseqkit/helper.go:86-103 */
  o->GapLetters = "- \t."; o->ValidateSeqLength = 10000; o->MaxLen = -1; o->MinLen = -1; /* bigseqkit/seq.go:32-55 */
  o->QualAsciiBase = 33; o->MinQual = -1; o->MaxQual = -1;
  o->FqEncoding = "sanger";                                                  /* bigseqkit/stats.go:28-38 */
  o->TranslTable = 1; o->Frame = "1";                                        /* bigseqkit/translate.go:22-35 */
  o->Region = "";
}

/* ---------------------------------------------------------------- alphabets
 * UNVERIFIED restatement of shenwei356/bio v0.7.0 seq/alphabet.go. */
enum { AB_NIL = 0, AB_DNA, AB_DNARED, AB_RNA, AB_RNARED, AB_PROTEIN, AB_UNLIMIT };
static const char *ab_name[] = {"", "DNA", "DNAredundant", "RNA", "RNAredundant", "Protein", "Unlimit"};
static uint8_t ab_valid[7][256], ab_pair[7][256];
static int ab_ready = 0;
static void ab_def(int a, const char *letters, const char *pairs, const char *gap, const char *amb) {
  for (int i = 0; i < 256; i++) ab_pair[a][i] = (uint8_t)i; /* PairLetter: unknown byte -> itself */
  for (size_t i = 0; letters[i]; i++) {
    ab_valid[a][(uint8_t)letters[i]] = 1;
    ab_pair[a][(uint8_t)letters[i]] = (uint8_t)pairs[i];
  }
  for (size_t i = 0; gap[i]; i++) ab_valid[a][(uint8_t)gap[i]] = 1;
  for (size_t i = 0; amb[i]; i++) ab_valid[a][(uint8_t)amb[i]] = 1;
}
static void ab_init(void) {
  if (ab_ready) return;
  ab_def(AB_DNA, "acgtACGT", "tgcaTGCA", " -.", "nN.");
  ab_def(AB_DNARED, "acgtryswkmbdhvACGTRYSWKMBDHV", "tgcayrswmkvhdbTGCAYRSWMKVHDB", " -.", "nN.");
  ab_def(AB_RNA, "acguACGU", "ugcaUGCA", " -.", "nN.");
  ab_def(AB_RNARED, "acguryswkmbdhvACGURYSWKMBDHV", "ugcayrswmkvhdbUGCAYRSWMKVHDB", " -.", "nN.");
  ab_def(AB_PROTEIN, "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ*_.",
         "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ*_.", " -", "xXbBzZ");
  for (int i = 0; i < 256; i++) { ab_valid[AB_UNLIMIT][i] = 1; ab_pair[AB_UNLIMIT][i] = (uint8_t)i; ab_pair[AB_NIL][i] = (uint8_t)i; }
  ab_ready = 1;
}
static int ab_is_valid(int a, const uint8_t *s, size_t n) {
  for (size_t i = 0; i < n; i++) if (!ab_valid[a][s[i]]) return 0;
  return 1;
}
/* seq.GuessAlphabetLessConservatively (call site lib/helper.go:288) */
static int ab_guess(const uint8_t *s, size_t n, int threshold) {
  if (n == 0) return AB_UNLIMIT;
  if (threshold > 0 && n > (size_t)threshold) n = (size_t)threshold;
  static const int order[] = {AB_DNA, AB_RNA, AB_DNARED, AB_RNARED, AB_PROTEIN};
  for (int k = 0; k < 5; k++)
    if (ab_is_valid(order[k], s, n)) {
      int a = order[k];
      if (a == AB_DNA) return AB_DNARED;
      if (a == AB_RNA) return AB_RNARED;
      return a;
    }
  return AB_UNLIMIT;
}
/* KitConfig.GetAlphabet (bigseqkit/helper.go:68-84) */
static int ab_from_type(const char *t, char *err) {
  if (!t || !strcasecmp(t, "auto")) return AB_NIL;
  if (!strcasecmp(t, "dna")) return AB_DNARED;
  if (!strcasecmp(t, "rna")) return AB_RNARED;
  if (!strcasecmp(t, "protein")) return AB_PROTEIN;
  if (!strcasecmp(t, "unlimit")) return AB_UNLIMIT;
  snprintf(err, 512, "invalid sequence type: %s, available value: dna|rna|protein|unlimit|auto", t);
  return -1;
}

/* ------------------------------------------------------------------ framing
 * worker.PlainFile(path, delim) + ReadFixer (bigseqkit/helper.go:148-178,
 * lib/helper.go:41-66).  IgnisHPC is not in the tree; pinned rule:
 *   FASTA: a record starts at BOF or at '>' right after '\n'.
 *   FASTQ ("\n@!\n+"): a record starts at BOF or at '@' right after '\n',
 *          unless the two bytes before that '\n' are "\n+".
 * Element = record text minus ONE trailing '\n'; the format is FASTQ iff the
 * first byte is '@' (lib/helper.go:228-230). */
size_t orc_frame(const uint8_t *d, size_t n, uint64_t **starts_out) {
  size_t cap = 1024, cnt = 0;
  uint64_t *st = (uint64_t *)malloc(cap * sizeof(uint64_t));
  if (n > 0) {
    int fq = d[0] == '@';
    st[cnt++] = 0;
    for (size_t i = 0; i + 1 < n; i++) {
      if (d[i] != '\n') continue;
      int hit;
      if (fq) hit = d[i + 1] == '@' && !(i >= 2 && d[i - 2] == '\n' && d[i - 1] == '+');
      else hit = d[i + 1] == '>';
      if (hit) {
        if (cnt + 2 > cap) { cap *= 2; st = (uint64_t *)realloc(st, cap * sizeof(uint64_t)); }
        st[cnt++] = i + 1;
      }
    }
  }
  st[cnt] = n;
  *starts_out = st;
  return cnt;
}

/* ------------------------------------------------------------------- parser
 * SeqParser (lib/helper.go:160-376) with SURVEY Q1 applied: head excludes the
 * '>'/'@' marker.  The literal `len(head)==0 && len(seq)==0 -> EOF` test
 * (:293-295) can never fire in the reference because its head still carries
 * the marker, so it is not restated. */
typedef struct {
  const uint8_t *head; size_t head_len;
  const uint8_t *id; size_t id_len;
  const uint8_t *desc; size_t desc_len;
  uint8_t *seq; size_t seq_len;
  uint8_t *qual; size_t qual_len;
} rec_t;

typedef struct {
  const uint8_t *d; size_t n;
  uint64_t *starts; size_t n_rec, cur;
  int is_fastq, alphabet, firstseq, id_ncbi, guess_len;
  int validate, validate_len;
  buf_t seqb, qualb;
  rec_t r;
  char err[512];
} parser_t;

static void parser_init(parser_t *p, const uint8_t *d, size_t n, int alphabet, const orc_opts *o) {
  memset(p, 0, sizeof *p);
  ab_init();
  p->d = d; p->n = n;
  p->n_rec = orc_frame(d, n, &p->starts);
  p->alphabet = alphabet; p->firstseq = 1;
  p->id_ncbi = o->IDNCBI; p->guess_len = o->AlphabetGuessSeqLength;
  p->is_fastq = n > 0 && d[0] == '@';
}
static void parser_free(parser_t *p) { free(p->starts); free(p->seqb.p); free(p->qualb.p); }

/* parseHeadIDAndDesc (lib/helper.go:329-369), default regexp ^(\S+)\s? ; the
 * blank-skipping loop advances twice per blank (sic, :334-339). */
static void parse_id_desc(parser_t *p, rec_t *r) {
  const uint8_t *h = r->head; size_t e = r->head_len;
  r->id = h; r->id_len = e; r->desc = h + e; r->desc_len = 0;
  if (p->id_ncbi) { /* regexp `\|([^\|]+)\| ` (bigseqkit/helper.go:97-100): leftmost match */
    for (size_t i = 0; i < e; i++) {
      if (h[i] != '|') continue;
      size_t j = i + 1;
      while (j < e && h[j] != '|') j++;
      if (j < e && j > i + 1 && j + 1 < e && h[j + 1] == ' ') { r->id = h + i + 1; r->id_len = j - i - 1; return; }
    }
    return; /* no match: ID = head, Desc = "" (:364-368) */
  }
  const uint8_t *sp = (const uint8_t *)memchr(h, ' ', e);
  size_t i = sp ? (size_t)(sp - h) : 0;
  if (!(sp && i > 0)) {
    sp = (const uint8_t *)memchr(h, '\t', e);
    i = sp ? (size_t)(sp - h) : 0;
    if (!(sp && i > 0)) return;
  }
  size_t j = i + 1;
  for (; j < e; j++) {
    if (h[j] == ' ' || h[j] == '\t') j++;
    else break;
  }
  r->id_len = i;
  if (j >= e) { r->desc = h + e; r->desc_len = 0; }
  else { r->desc = h + j; r->desc_len = e - j; }
}

/* SeqParser.Read (lib/helper.go:219-325).  Returns 1 record, 0 EOF, -1 error. */
static int parser_read(parser_t *p) {
  if (p->cur >= p->n_rec) return 0;
  size_t s = p->starts[p->cur], e = p->starts[p->cur + 1];
  p->cur++;
  /* ReadFixer: strip one trailing '\n' (lib/helper.go:51) */
  if (e > s && p->d[e - 1] == '\n') e--;
  const uint8_t *b = p->d + s; size_t m = e - s;
  char marker = p->is_fastq ? '@' : '>';
  if (m > 0 && b[0] == (uint8_t)marker) { b++; m--; } /* else: ReadFixer prepends the marker (:52-61) */
  rec_t *r = &p->r;
  p->seqb.n = 0; p->qualb.n = 0;
  const uint8_t *nl = (const uint8_t *)memchr(b, '\n', m);
  if (nl) {
    r->head = b; r->head_len = (size_t)(nl - b);
    const uint8_t *q = nl + 1, *end = b + m;
    if (!p->is_fastq) { /* :240-250 */
      for (;;) {
        const uint8_t *k = (const uint8_t *)memchr(q, '\n', (size_t)(end - q));
        if (k) { buf_add(&p->seqb, q, (size_t)(k - q)); q = k + 1; continue; }
        buf_add(&p->seqb, q, (size_t)(end - q));
        break;
      }
    } else { /* :251-273 */
      int is_qual = 0;
      for (;;) {
        const uint8_t *k = (const uint8_t *)memchr(q, '\n', (size_t)(end - q));
        if (k) {
          size_t len = (size_t)(k - q);
          if (len > 0 && q[0] == '+' && !is_qual) is_qual = 1;
          else if (is_qual) buf_add(&p->qualb, q, len);
          else buf_add(&p->seqb, q, len);
          q = k + 1;
          continue;
        }
        if (is_qual) buf_add(&p->qualb, q, (size_t)(end - q));
        break;
      }
    }
  } else { /* :275-283 */
    r->head = b; r->head_len = m;
  }
  buf_reserve(&p->seqb, 1); buf_reserve(&p->qualb, 1);
  r->seq = p->seqb.p; r->seq_len = p->seqb.n;
  r->qual = p->qualb.p; r->qual_len = p->is_fastq ? p->qualb.n : 0;
  if (p->firstseq) { /* :286-291 */
    if (p->alphabet == AB_NIL) p->alphabet = ab_guess(r->seq, r->seq_len, p->guess_len);
    p->firstseq = 0;
  }
  parse_id_desc(p, r);
  if (p->validate) { /* :304-306,319-321 ; bio Alphabet.IsValid on the first ValidSeqLengthThreshold bytes */
    size_t l = r->seq_len;
    if (p->validate_len > 0 && l > (size_t)p->validate_len) l = (size_t)p->validate_len;
    for (size_t i = 0; i < l; i++)
      if (!ab_valid[p->alphabet][r->seq[i]]) {
        snprintf(p->err, sizeof p->err, "seq: invalid %s letter: %c", ab_name[p->alphabet], r->seq[i]);
        return -1;
      }
  }
  if (p->is_fastq && r->seq_len != r->qual_len) { /* :308-311 */
    snprintf(p->err, sizeof p->err, "seq('%.*s'): unmatched length of sequence (%zu) and quality (%zu)",
             (int)r->head_len, (const char *)r->head, r->seq_len, r->qual_len);
    return -1;
  }
  return 1;
}
static int parser_alphabet(parser_t *p) { return p->alphabet == AB_NIL ? AB_UNLIMIT : p->alphabet; } /* :371-376 */

/* -------------------------------------------------------- bio leaf helpers */
/* wrapByteSlice (lib/helper.go:81-117) == byteutil.WrapByteSlice */
size_t orc_wrap_len(size_t l, int width) {
  if (width < 1 || l == 0) return l;
  size_t lines = (l % (size_t)width == 0) ? l / (size_t)width - 1 : l / (size_t)width;
  return l + lines;
}
static void wrap_into(buf_t *b, const uint8_t *s, size_t l, int width) {
  if (width < 1 || l == 0) { buf_add(b, s, l); return; }
  size_t w = (size_t)width;
  size_t lines = (l % w == 0) ? l / w - 1 : l / w;
  for (size_t i = 0; i <= lines; i++) {
    size_t st = i * w, en = (i + 1) * w;
    if (en > l) en = l;
    buf_add(b, s + st, en - st);
    if (i < lines) buf_addc(b, '\n');
  }
}
static void reverse_bytes(uint8_t *s, size_t n) {
  if (n < 2) return;
  for (size_t i = 0, j = n - 1; i < j; i++, j--) { uint8_t t = s[i]; s[i] = s[j]; s[j] = t; }
}
/* Seq.ComplementInplace: no-op for Unlimit; PairLetter per byte (UNVERIFIED bio) */
static void complement_bytes(int ab, uint8_t *s, size_t n) {
  if (ab == AB_UNLIMIT || ab == AB_NIL) return;
  for (size_t i = 0; i < n; i++) s[i] = ab_pair[ab][s[i]];
}
/* Seq.RevCom: copy -> reverse -> complement */
static void revcom_into(buf_t *b, int ab, const uint8_t *s, size_t n) {
  b->n = 0; buf_reserve(b, n + 1);
  for (size_t i = 0; i < n; i++) b->p[i] = s[n - 1 - i];
  b->n = n;
  complement_bytes(ab, b->p, n);
}
/* seq.SubLocation (UNVERIFIED bio; pinned by cli/helper.go:348-361) */
size_t orc_subseq_range(size_t length, int start, int end, size_t *s0) {
  long len = (long)length, s = start, e = end;
  *s0 = 0;
  if (len == 0) return 0;
  if (s < 1) {
    if (s == 0) s = 1;
    else {
      if (e < 0 && s > e) return 0;
      if (-s > len) s = 1; else s = len + s + 1;
    }
  } else if (s > len) return 0;
  if (e > len) e = len;
  else if (e < 1) {
    if (e == 0) e = -1;
    if (-e > len) return 0;
    e = len + e + 1;
  }
  if (s - 1 > e) return 0;
  *s0 = (size_t)(s - 1);
  return (size_t)(e - (s - 1));
}
/* fastx.Record.Format(width) (UNVERIFIED bio): FASTQ when qual non-empty or ForcelyOutputFastq */
static void format_record(buf_t *b, const uint8_t *name, size_t name_len, const uint8_t *seq, size_t seq_len,
                          const uint8_t *qual, size_t qual_len, int fastq, int width) {
  b->n = 0;
  if (qual_len > 0 || fastq) {
    buf_addc(b, '@'); buf_add(b, name, name_len); buf_addc(b, '\n');
    wrap_into(b, seq, seq_len, width);
    buf_adds(b, "\n+\n");
    wrap_into(b, qual, qual_len, width);
    buf_addc(b, '\n');
  } else {
    buf_addc(b, '>'); buf_add(b, name, name_len); buf_addc(b, '\n');
    wrap_into(b, seq, seq_len, width);
    buf_addc(b, '\n');
  }
}
static void lower_bytes(uint8_t *s, size_t n) { for (size_t i = 0; i < n; i++) if (s[i] >= 'A' && s[i] <= 'Z') s[i] += 32; }
static void upper_bytes(uint8_t *s, size_t n) { for (size_t i = 0; i < n; i++) if (s[i] >= 'a' && s[i] <= 'z') s[i] -= 32; }

/* ------------------------------------------------------------------- XXH64
 * cespare/xxhash v2 Sum64 == XXH64 seed 0 (call sites lib/rmdup.go:67-85). */
#define P1 11400714785074694791ULL
#define P2 14029467366897019727ULL
#define P3 1609587929392839161ULL
#define P4 9650029242287828579ULL
#define P5 2870177450012600261ULL
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t xxround(uint64_t acc, uint64_t in) { acc += in * P2; acc = rotl64(acc, 31); return acc * P1; }
static inline uint64_t xxmerge(uint64_t acc, uint64_t v) { v = xxround(0, v); acc ^= v; return acc * P1 + P4; }
uint64_t orc_xxh64(const uint8_t *p, size_t len, uint64_t seed) {
  const uint8_t *end = p + len;
  uint64_t h;
  if (len >= 32) {
    const uint8_t *lim = end - 32;
    uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    do {
      v1 = xxround(v1, rd64(p)); v2 = xxround(v2, rd64(p + 8));
      v3 = xxround(v3, rd64(p + 16)); v4 = xxround(v4, rd64(p + 24));
      p += 32;
    } while (p <= lim);
    h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
    h = xxmerge(h, v1); h = xxmerge(h, v2); h = xxmerge(h, v3); h = xxmerge(h, v4);
  } else h = seed + P5;
  h += (uint64_t)len;
  while (p + 8 <= end) { h ^= xxround(0, rd64(p)); h = rotl64(h, 27) * P1 + P4; p += 8; }
  if (p + 4 <= end) { h ^= (uint64_t)rd32(p) * P1; h = rotl64(h, 23) * P2 + P3; p += 4; }
  while (p < end) { h ^= (uint64_t)(*p) * P5; h = rotl64(h, 11) * P1; p++; }
  h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
  return h;
}

/* --------------------------------------------------------------------- seq
 * SeqTransform.Before/Call (lib/seq.go:28-269). */
int orc_seq(const uint8_t *data, size_t n, const orc_opts *o, orc_out *out) {
  memset(out, 0, sizeof *out);
  int ab = ab_from_type(o->SeqType, out->err);
  if (ab < 0) return -1;
  if (!o->GapLetters || !o->GapLetters[0]) { snprintf(out->err, 512, "value of flag -G (--gap-letters) should not be empty"); return -1; }
  for (const char *c = o->GapLetters; *c; c++)
    if ((unsigned char)*c > 127) { snprintf(out->err, 512, "value of -G (--gap-letters) contains non-ASCII characters"); return -1; }
  if (o->MinLen >= 0 && o->MaxLen >= 0 && o->MinLen > o->MaxLen) { snprintf(out->err, 512, "value of flag -m (--min-len) should be >= value of flag -M (--max-len)"); return -1; }
  if (o->MinQual >= 0 && o->MaxQual >= 0 && o->MinQual > o->MaxQual) { snprintf(out->err, 512, "value of flag -Q (--min-qual) should be <= value of flag -R (--max-qual)"); return -1; }
  if (o->LowerCase && o->UpperCase) { snprintf(out->err, 512, "could not give both flags -l (--lower-case) and -u (--upper-case)"); return -1; }
  int validate = o->ValidateSeq;
  if (!validate && !(ab == AB_NIL || ab == AB_UNLIMIT)) validate = 1; /* :66-72 */

  parser_t p; parser_init(&p, data, n, ab, o);
  p.validate = validate; p.validate_len = o->ValidateSeqLength;
  sink_t sink; memset(&sink, 0, sizeof sink);
  buf_t ob; memset(&ob, 0, sizeof ob);
  uint8_t gap[256]; memset(gap, 0, sizeof gap);
  for (const char *c = o->GapLetters; *c; c++) gap[(uint8_t)*c] = 1;
  int filterMinLen = o->MinLen > 0, filterMaxLen = o->MaxLen > 0;
  int filterMinQual = o->MinQual > 0, filterMaxQual = o->MaxQual > 0;
  int line_width = o->LineWidth;
  if (o->Seq || o->Qual) line_width = 0; /* :106-108 */
  int check_type = 1, is_fastq = 0, print_qual = 0, rc;
  while ((rc = parser_read(&p)) > 0) {
    rec_t *r = &p.r;
    ob.n = 0;
    if (check_type) { is_fastq = p.is_fastq; if (is_fastq) { line_width = 0; print_qual = 1; } check_type = 0; }
    if (o->RemoveGaps) { /* Seq.RemoveGapsInplace: seq and the same positions of qual */
      size_t w = 0;
      for (size_t i = 0; i < r->seq_len; i++)
        if (!gap[r->seq[i]]) { r->seq[w] = r->seq[i]; if (r->qual_len) r->qual[w] = r->qual[i]; w++; }
      r->seq_len = w; if (r->qual_len) r->qual_len = w;
    }
    if (filterMinLen && (long)r->seq_len < o->MinLen) continue;
    if (filterMaxLen && (long)r->seq_len > o->MaxLen) continue;
    if (filterMinQual || filterMaxQual) { /* Seq.AvgQual(base) (UNVERIFIED bio) */
      double avg = 0;
      if (r->qual_len > 0) {
        double sum = 0;
        for (size_t i = 0; i < r->qual_len; i++) sum += pow(10, (double)((int)r->qual[i] - o->QualAsciiBase) / -10);
        avg = -10 * log10(sum / (double)r->qual_len);
      }
      if (filterMinQual && avg < o->MinQual) continue;
      if (filterMaxQual && avg >= o->MaxQual) continue;
    }
    int print_name = 1, print_seq = 1;
    if (o->Name && o->Seq) { }
    else if (o->Name) { print_name = 1; print_seq = 0; print_qual = 0; }
    else if (o->Seq) { print_name = 0; print_seq = 1; print_qual = 0; }
    else if (o->Qual) {
      if (!is_fastq) { snprintf(out->err, 512, "FASTA format has no quality. So do not just use flag -q (--qual)"); rc = -2; break; }
      print_name = 0; print_seq = 0; print_qual = 1;
    }
    if (print_name) {
      const uint8_t *h = o->OnlyId ? r->id : r->head; size_t hl = o->OnlyId ? r->id_len : r->head_len;
      if (print_seq) buf_addc(&ob, is_fastq ? '@' : '>');
      buf_add(&ob, h, hl); buf_addc(&ob, '\n');
    }
    if (o->Reverse) { reverse_bytes(r->seq, r->seq_len); if (r->qual_len) reverse_bytes(r->qual, r->qual_len); }
    if (o->Complement) complement_bytes(p.alphabet, r->seq, r->seq_len);
    if (print_seq) {
      int a = parser_alphabet(&p);
      if (o->Dna2rna && !(a == AB_RNA || a == AB_RNARED))
        for (size_t i = 0; i < r->seq_len; i++) { if (r->seq[i] == 't') r->seq[i] = 'u'; else if (r->seq[i] == 'T') r->seq[i] = 'U'; }
      if (o->Rna2dna && !(a == AB_DNA || a == AB_DNARED))
        for (size_t i = 0; i < r->seq_len; i++) { if (r->seq[i] == 'u') r->seq[i] = 't'; else if (r->seq[i] == 'U') r->seq[i] = 'T'; }
      if (o->LowerCase) lower_bytes(r->seq, r->seq_len);
      else if (o->UpperCase) upper_bytes(r->seq, r->seq_len);
      if (is_fastq) buf_add(&ob, r->seq, r->seq_len);
      else wrap_into(&ob, r->seq, r->seq_len, line_width);
      buf_addc(&ob, '\n');
    }
    if (print_qual) {
      if (!o->Qual) buf_adds(&ob, "+\n");
      buf_add(&ob, r->qual, r->qual_len);
      buf_addc(&ob, '\n');
    }
    size_t l = ob.n;
    if (l && ob.p[l - 1] == '\n') l--; /* :261-264 */
    sink_elem(&sink, ob.p, l);
  }
  if (rc == -1) snprintf(out->err, 512, "%s", p.err);
  sink_to_out(&sink, out);
  free(ob.p); parser_free(&p);
  return rc < 0 ? -1 : 0;
}

/* ------------------------------------------------------------------- stats
 * Stats.Call (lib/stats.go:48-117) + Stats finalise (bigseqkit/stats.go:75-166). */
typedef struct { uint64_t *k, *v; size_t cap, n; } hmap_t; /* open addressing; key+1 stored, 0 = empty */
static void hmap_grow(hmap_t *m);
static void hmap_inc(hmap_t *m, uint64_t key, uint64_t by) {
  if ((m->n + 1) * 2 > m->cap) hmap_grow(m);
  size_t mask = m->cap - 1, i = (size_t)(key * 0x9E3779B97F4A7C15ULL >> 20) & mask;
  for (;;) {
    if (m->k[i] == 0) { m->k[i] = key + 1; m->v[i] = by; m->n++; return; }
    if (m->k[i] == key + 1) { m->v[i] += by; return; }
    i = (i + 1) & mask;
  }
}
static void hmap_grow(hmap_t *m) {
  hmap_t o = *m;
  m->cap = o.cap ? o.cap * 2 : 64; m->n = 0;
  m->k = (uint64_t *)calloc(m->cap, 8); m->v = (uint64_t *)calloc(m->cap, 8);
  for (size_t i = 0; i < o.cap; i++) if (o.k[i]) hmap_inc(m, o.k[i] - 1, o.v[i]);
  free(o.k); free(o.v);
}
static int cmp_u64pair(const void *a, const void *b) {
  uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return x < y ? -1 : x > y;
}
static void hmap_to_sorted(hmap_t *m, orc_stats *s) {
  uint64_t *pairs = (uint64_t *)malloc((m->n + 1) * 16);
  size_t c = 0;
  for (size_t i = 0; i < m->cap; i++) if (m->k[i]) { pairs[2 * c] = m->k[i] - 1; pairs[2 * c + 1] = m->v[i]; c++; }
  qsort(pairs, c, 16, cmp_u64pair);
  s->hist_len = (uint64_t *)malloc((c + 1) * 8); s->hist_cnt = (uint64_t *)malloc((c + 1) * 8);
  for (size_t i = 0; i < c; i++) { s->hist_len[i] = pairs[2 * i]; s->hist_cnt[i] = pairs[2 * i + 1]; }
  s->n_hist = c;
  free(pairs);
}
static int fq_offset(const char *enc, char *err) { /* parseQualityEncoding (lib/helper.go:119-136) + QualityEncoding.Offset */
  if (!enc || !*enc) return 0;
  if (!strcasecmp(enc, "sanger") || !strcasecmp(enc, "illumina-1.8+")) return 33;
  if (!strcasecmp(enc, "solexa") || !strcasecmp(enc, "illumina-1.3+") || !strcasecmp(enc, "illumina-1.5+")) return 64;
  snprintf(err, 512, "unsupported quality encoding: %s. available values: 'sanger', 'solexa', 'illumina-1.3+', 'illumina-1.5+', 'illumina-1.8+'", enc);
  return -1;
}
static int stats_partial(const uint8_t *data, size_t n, const orc_opts *o, hmap_t *hist, uint64_t *q20, uint64_t *q30,
                         uint64_t *gaps, char *type, uint64_t *nrec, char *err) {
  int ab = ab_from_type(o->SeqType, err);
  if (ab < 0) return -1;
  const char *gl = o->GapLetters ? o->GapLetters : "- .";
  if (!gl[0]) { snprintf(err, 512, "value of flag -G (--gap-letters) should not be empty"); return -1; }
  int off = fq_offset(o->FqEncoding, err);
  if (off < 0) return -1;
  uint8_t gap[256]; memset(gap, 0, sizeof gap);
  for (const char *c = gl; *c; c++) gap[(uint8_t)*c] = 1;
  parser_t p; parser_init(&p, data, n, ab, o);
  int rc, any = 0;
  const uint8_t *first_seq = NULL; size_t first_len = 0; buf_t fs; memset(&fs, 0, sizeof fs);
  while ((rc = parser_read(&p)) > 0) {
    rec_t *r = &p.r;
    if (!any) { buf_add(&fs, r->seq, r->seq_len); first_seq = fs.p; first_len = fs.n; any = 1; }
    hmap_inc(hist, r->seq_len, 1); /* :88 */
    if (o->All) {
      if (p.is_fastq)
        for (size_t i = 0; i < r->qual_len; i++) {
          int q = (int)r->qual[i] - off;
          if (q >= 20) { (*q20)++; if (q >= 30) (*q30)++; }
        }
      uint64_t g = 0;
      for (size_t i = 0; i < r->seq_len; i++) g += gap[r->seq[i]];
      *gaps += g;
    }
    (*nrec)++;
  }
  if (rc < 0) { snprintf(err, 512, "%s", p.err); parser_free(&p); free(fs.p); return -1; }
  /* type tag (lib/stats.go:106-114) then bigseqkit/stats.go:109-130 */
  int a = parser_alphabet(&p);
  if (a == AB_DNARED) strcpy(type, "DNA");
  else if (a == AB_RNARED) strcpy(type, "RNA");
  else if (!any && a == AB_UNLIMIT) strcpy(type, "");
  else strcpy(type, ab_name[ab_guess(first_seq, first_len, o->AlphabetGuessSeqLength)]); /* bio reader on Take(1) */
  parser_free(&p); free(fs.p);
  return 0;
}
/* util/math.Round (UNVERIFIED shenwei356/util v0.5.0) */
static double round_n(double f, int n) { double p = pow(10, n); return trunc((f + 0.5 / p) * p) / p; }
static uint64_t hist_at(const orc_stats *s, uint64_t idx) { /* idx-th smallest length (0-based) */
  uint64_t c = 0;
  for (size_t i = 0; i < s->n_hist; i++) { c += s->hist_cnt[i]; if (idx < c) return s->hist_len[i]; }
  return 0;
}
static double hist_val(const orc_stats *s, int even, uint64_t l, uint64_t r) {
  if (even) return ((double)hist_at(s, l) + (double)hist_at(s, r)) / 2;
  return (double)hist_at(s, l);
}
/* bio/util.LengthStats (UNVERIFIED) + bigseqkit/stats.go:132-161 */
void orc_stats_finalise(orc_stats *s, int all) {
  uint64_t num = 0, sum = 0;
  for (size_t i = 0; i < s->n_hist; i++) { num += s->hist_cnt[i]; sum += s->hist_len[i] * s->hist_cnt[i]; }
  s->num = num; s->sum_len = sum;
  s->min_len = s->max_len = 0; s->avg_len = 0; s->n50 = s->l50 = 0; s->q1 = s->q2 = s->q3 = 0; s->q20_pct = s->q30_pct = 0;
  if (num == 0) { s->sum_gap = 0; return; } /* :149-161: all zeros */
  s->min_len = s->hist_len[0]; s->max_len = s->hist_len[s->n_hist - 1];
  s->avg_len = round_n((double)sum / (double)num, 1);
  if (all) {
    if (s->n_hist == 1) { s->n50 = s->hist_len[0]; s->q1 = s->q2 = s->q3 = (double)s->hist_len[0]; }
    else {
      double acc = 0, half = (double)sum / 2;
      uint64_t l50 = 0;
      for (size_t i = s->n_hist; i-- > 0;) {
        acc += (double)(s->hist_len[i] * s->hist_cnt[i]); l50 += s->hist_cnt[i];
        if (acc >= half) { s->n50 = s->hist_len[i]; s->l50 = l50; break; }
      }
      int even = (num & 1) == 0;
      s->q2 = even ? hist_val(s, 1, num / 2 - 1, num / 2) : hist_val(s, 0, num / 2, 0);
      uint64_t h = even ? num / 2 : (num + 1) / 2, mean = num / 2;
      int heven = (h % 2) == 0;
      s->q1 = heven ? hist_val(s, 1, h / 2 - 1, h / 2) : hist_val(s, 0, h / 2, 0);
      s->q3 = heven ? hist_val(s, 1, mean + h / 2 - 1, mean + h / 2) : hist_val(s, 0, mean + h / 2, 0);
    }
  }
  s->q20_pct = round_n((double)s->q20 / (double)sum * 100, 2);
  s->q30_pct = round_n((double)s->q30 / (double)sum * 100, 2);
}
int orc_stats_run(const uint8_t *data, size_t n, const orc_opts *o, orc_stats *s) {
  memset(s, 0, sizeof *s);
  hmap_t h; memset(&h, 0, sizeof h);
  uint64_t nrec = 0;
  int rc = stats_partial(data, n, o, &h, &s->q20, &s->q30, &s->sum_gap, s->type, &nrec, s->err);
  if (rc == 0) { hmap_to_sorted(&h, s); orc_stats_finalise(s, o->All); }
  free(h.k); free(h.v);
  return rc;
}
/* StatsReduce with sum semantics (literal lib/stats.go:128-137 overwrites) */
int orc_stats_merge(orc_stats *d, const orc_stats *src) {
  hmap_t h; memset(&h, 0, sizeof h);
  for (size_t i = 0; i < d->n_hist; i++) hmap_inc(&h, d->hist_len[i], d->hist_cnt[i]);
  for (size_t i = 0; i < src->n_hist; i++) hmap_inc(&h, src->hist_len[i], src->hist_cnt[i]);
  free(d->hist_len); free(d->hist_cnt);
  hmap_to_sorted(&h, d);
  free(h.k); free(h.v);
  d->q20 += src->q20; d->q30 += src->q30; d->sum_gap += src->sum_gap;
  if (!d->type[0]) strcpy(d->type, src->type);
  return 0;
}
/* humanize.Comma / Commaf (UNVERIFIED dustin/go-humanize) */
static void comma_u(char *dst, uint64_t v) {
  char t[32]; int n = snprintf(t, sizeof t, "%llu", (unsigned long long)v), k = 0;
  for (int i = 0; i < n; i++) { dst[k++] = t[i]; if ((n - 1 - i) % 3 == 0 && i != n - 1) dst[k++] = ','; }
  dst[k] = 0;
}
static void comma_f(char *dst, double v) {
  char t[64]; snprintf(t, sizeof t, "%.10g", v); /* strconv 'f' -1 shortest; our values have <=2 decimals */
  char *dot = strchr(t, '.');
  char ip[32]; size_t il = dot ? (size_t)(dot - t) : strlen(t);
  memcpy(ip, t, il); ip[il] = 0;
  comma_u(dst, strtoull(ip, NULL, 10));
  if (dot) strcat(dst, dot);
}
/* StatsString (bigseqkit/stats.go:168-288).  Pretty mode: tatsushid/go-prettytable
 * default separator " " (UNVERIFIED). */
char *orc_stats_render(const orc_stats *s, const char *file, const char *format, int tabular, int all) {
  buf_t b; memset(&b, 0, sizeof b);
  if (tabular) {
    buf_adds(&b, "file\tformat\ttype\tnum_seqs\tsum_len\tmin_len\tavg_len\tmax_len");
    if (all) buf_adds(&b, "\tQ1\tQ2\tQ3\tsum_gap\tN50\tQ20(%)\tQ30(%)");
    buf_addc(&b, '\n');
    buf_printf(&b, "%s\t%s\t%s\t%llu\t%llu\t%llu\t%.1f\t%llu", file, format, s->type, (unsigned long long)s->num,
               (unsigned long long)s->sum_len, (unsigned long long)s->min_len, s->avg_len, (unsigned long long)s->max_len);
    if (all)
      buf_printf(&b, "\t%.1f\t%.1f\t%.1f\t%llu\t%llu\t%.2f\t%.2f", s->q1, s->q2, s->q3, (unsigned long long)s->sum_gap,
                 (unsigned long long)s->n50, s->q20_pct, s->q30_pct);
    buf_addc(&b, '\n');
  } else {
    const char *hdr[15] = {"file", "format", "type", "num_seqs", "sum_len", "min_len", "avg_len", "max_len",
                           "Q1", "Q2", "Q3", "sum_gap", "N50", "Q20(%)", "Q30(%)"};
    char cell[15][64]; int nc = all ? 15 : 8;
    snprintf(cell[0], 64, "%s", file); snprintf(cell[1], 64, "%s", format); snprintf(cell[2], 64, "%s", s->type);
    comma_u(cell[3], s->num); comma_u(cell[4], s->sum_len); comma_u(cell[5], s->min_len);
    comma_f(cell[6], s->avg_len); comma_u(cell[7], s->max_len);
    if (all) {
      comma_f(cell[8], s->q1); comma_f(cell[9], s->q2); comma_f(cell[10], s->q3);
      comma_u(cell[11], s->sum_gap); comma_u(cell[12], s->n50); comma_f(cell[13], s->q20_pct); comma_f(cell[14], s->q30_pct);
    }
    for (int row = 0; row < 2; row++) {
      for (int c = 0; c < nc; c++) {
        const char *txt = row == 0 ? hdr[c] : cell[c];
        size_t w = strlen(hdr[c]) > strlen(cell[c]) ? strlen(hdr[c]) : strlen(cell[c]);
        size_t pad = w - strlen(txt);
        int right = c >= 3;
        if (c) buf_addc(&b, ' ');
        if (right) for (size_t i = 0; i < pad; i++) buf_addc(&b, ' ');
        buf_adds(&b, txt);
        if (!right) for (size_t i = 0; i < pad; i++) buf_addc(&b, ' ');
      }
      buf_addc(&b, '\n');
    }
  }
  buf_addc(&b, 0);
  return (char *)b.p;
}

/* ------------------------------------------------------------------- rmdup
 * RmDupPrepare + GroupByKey + RmDupCheck (lib/rmdup.go:43-242, bigseqkit/rmdup.go:70-108)
 * with SURVEY Q4: first occurrence in input order wins, output in input order,
 * exact-subject compare inside an equal-key group. */
static int rmdup_check_flags(const orc_opts *o, char *err) { /* bigseqkit/rmdup.go:79-85 */
  if (o->BySeq && o->ByName) { snprintf(err, 512, "only one/none of the flags -s (--by-seq) and -n (--by-name) is allowed"); return -1; }
  if (o->OnlyPositiveStrand && !o->BySeq) { snprintf(err, 512, "flag -s (--by-seq) needed when using -P (--only-positive-strand)"); return -1; }
  return 0;
}
static void rmdup_subject(const orc_opts *o, const rec_t *r, buf_t *tmp, const uint8_t **s, size_t *n) {
  const uint8_t *p; size_t l;
  if (o->BySeq) { p = r->seq; l = r->seq_len; }
  else if (o->ByName) { p = r->head; l = r->head_len; }
  else { p = r->id; l = r->id_len; }
  if (o->IgnoreCase) { tmp->n = 0; buf_add(tmp, p, l); lower_bytes(tmp->p, l); p = tmp->p; } /* bytes.ToLower */
  *s = p; *n = l;
}
int orc_rmdup_keys(const uint8_t *data, size_t n, const orc_opts *o, int64_t **keys, size_t *n_keys) {
  char err[512];
  int ab = ab_from_type(o->SeqType, err);
  if (ab < 0) return -1;
  parser_t p; parser_init(&p, data, n, ab, o);
  int64_t *k = (int64_t *)malloc((p.n_rec + 1) * 8);
  size_t c = 0; int rc; buf_t tmp; memset(&tmp, 0, sizeof tmp);
  while ((rc = parser_read(&p)) > 0) {
    const uint8_t *s; size_t l;
    rmdup_subject(o, &p.r, &tmp, &s, &l);
    k[c++] = (int64_t)orc_xxh64(s, l, 0); /* lib/rmdup.go:67-86 */
  }
  parser_free(&p); free(tmp.p);
  *keys = k; *n_keys = c;
  return rc < 0 ? -1 : 0;
}
typedef struct { uint64_t key; size_t off, len; } dd_ent;
int orc_rmdup(const uint8_t *data, size_t n, const orc_opts *o, orc_out *out, uint64_t *n_removed) {
  memset(out, 0, sizeof *out);
  if (rmdup_check_flags(o, out->err)) return -1;
  int ab = ab_from_type(o->SeqType, out->err);
  if (ab < 0) return -1;
  parser_t p; parser_init(&p, data, n, ab, o);
  sink_t sink; memset(&sink, 0, sizeof sink);
  buf_t ob, tmp, arena; memset(&ob, 0, sizeof ob); memset(&tmp, 0, sizeof tmp); memset(&arena, 0, sizeof arena);
  size_t cap = 64; while (cap < 2 * p.n_rec + 2) cap *= 2;
  dd_ent *tab = (dd_ent *)calloc(cap, sizeof(dd_ent)); /* len+1 stored in .len; 0 = empty */
  uint64_t removed = 0; int rc;
  while ((rc = parser_read(&p)) > 0) {
    rec_t *r = &p.r;
    const uint8_t *s; size_t l;
    rmdup_subject(o, r, &tmp, &s, &l);
    uint64_t key = orc_xxh64(s, l, 0);
    size_t i = (size_t)(key & (cap - 1));
    int dup = 0;
    for (;;) {
      if (tab[i].len == 0) break;
      if (tab[i].key == key && tab[i].len - 1 == l && memcmp(arena.p + tab[i].off, s, l) == 0) { dup = 1; break; } /* :180 */
      i = (i + 1) & (cap - 1);
    }
    if (dup) { removed++; continue; }
    tab[i].key = key; tab[i].off = arena.n; tab[i].len = l + 1;
    buf_add(&arena, s, l);
    int fq = p.is_fastq;
    format_record(&ob, r->head, r->head_len, r->seq, r->seq_len, r->qual, r->qual_len, fq, fq ? 0 : o->LineWidth);
    sink_elem(&sink, ob.p, ob.n - 1); /* :214-215 */
  }
  if (rc < 0) snprintf(out->err, 512, "%s", p.err);
  if (n_removed) *n_removed = removed;
  sink_to_out(&sink, out);
  free(tab); free(ob.p); free(tmp.p); free(arena.p); parser_free(&p);
  return rc < 0 ? -1 : 0;
}

/* rmdup -d / -D side outputs (lib/rmdup.go:180-239, After :245-275).  SURVEY Q4/Q6: first occurrence in input
 * order is the group's kept member; dup_seqs = every removed record as Record.Format(LineWidth), input order;
 * dup_num = one row "count\tid1, id2, ..." per subject with more than one member, rows in order of the group's
 * first member (the reference iterates a Go map there: order unspecified), ids in input order. */
typedef struct { uint64_t key; size_t off, len, grp; } dg_ent;
typedef struct { uint64_t count; buf_t ids; } dg_grp;
int orc_rmdup_dups(const uint8_t *data, size_t n, const orc_opts *o, orc_out *dup_seqs, orc_out *dup_num) {
  memset(dup_seqs, 0, sizeof *dup_seqs); memset(dup_num, 0, sizeof *dup_num);
  if (rmdup_check_flags(o, dup_seqs->err)) return -1;
  int ab = ab_from_type(o->SeqType, dup_seqs->err);
  if (ab < 0) return -1;
  parser_t p; parser_init(&p, data, n, ab, o);
  sink_t s_seq, s_num; memset(&s_seq, 0, sizeof s_seq); memset(&s_num, 0, sizeof s_num);
  buf_t ob, tmp, arena; memset(&ob, 0, sizeof ob); memset(&tmp, 0, sizeof tmp); memset(&arena, 0, sizeof arena);
  size_t cap = 64; while (cap < 2 * p.n_rec + 2) cap *= 2;
  dg_ent *tab = (dg_ent *)calloc(cap, sizeof(dg_ent));
  dg_grp *grp = (dg_grp *)calloc(p.n_rec + 1, sizeof(dg_grp));
  size_t n_grp = 0; int rc;
  while ((rc = parser_read(&p)) > 0) {
    rec_t *r = &p.r;
    const uint8_t *s; size_t l;
    rmdup_subject(o, r, &tmp, &s, &l);
    uint64_t key = orc_xxh64(s, l, 0);
    size_t i = (size_t)(key & (cap - 1));
    int dup = 0;
    for (;;) {
      if (tab[i].len == 0) break;
      if (tab[i].key == key && tab[i].len - 1 == l && memcmp(arena.p + tab[i].off, s, l) == 0) { dup = 1; break; }
      i = (i + 1) & (cap - 1);
    }
    if (dup) {
      dg_grp *g = &grp[tab[i].grp];
      g->count++;
      buf_add(&g->ids, (const uint8_t *)", ", 2); buf_add(&g->ids, r->id, r->id_len);       /* :188-190 */
      int fq = p.is_fastq;
      format_record(&ob, r->head, r->head_len, r->seq, r->seq_len, r->qual, r->qual_len, fq, fq ? 0 : o->LineWidth);
      sink_elem(&s_seq, ob.p, ob.n - 1);                                                      /* :185-187 */
      continue;
    }
    tab[i].key = key; tab[i].off = arena.n; tab[i].len = l + 1; tab[i].grp = n_grp;
    buf_add(&arena, s, l);
    grp[n_grp].count = 1; buf_add(&grp[n_grp].ids, r->id, r->id_len);                         /* :217-219 */
    n_grp++;
  }
  for (size_t g = 0; g < n_grp; g++) {
    if (grp[g].count > 1) {                                                                   /* :230-234 */
      char num[32]; int k = snprintf(num, sizeof num, "%llu\t", (unsigned long long)grp[g].count);
      ob.n = 0; buf_add(&ob, (const uint8_t *)num, (size_t)k); buf_add(&ob, grp[g].ids.p, grp[g].ids.n);
      sink_elem(&s_num, ob.p, ob.n);
    }
    free(grp[g].ids.p);
  }
  if (rc < 0) snprintf(dup_seqs->err, 512, "%s", p.err);
  sink_to_out(&s_seq, dup_seqs); sink_to_out(&s_num, dup_num);
  free(tab); free(grp); free(ob.p); free(tmp.p); free(arena.p); parser_free(&p);
  return rc < 0 ? -1 : 0;
}

/* --------------------------------------------------------------- translate
 * seq.CodonTables / Seq.Translate (UNVERIFIED bio v0.7.0 seq/codon_table.go;
 * NCBI gc.prt strings, base order TCAG) + Translate.Call (lib/translate.go:66-145)
 * with SURVEY Q3: one element per (record, frame). */
typedef struct { int id; const char *name, *aas, *starts; } gcode_t;
static const gcode_t gcodes[] = {
  {1, "The Standard Code", "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "---M------**--*----M---------------M----------------------------"},
  {2, "The Vertebrate Mitochondrial Code", "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSS**VVVVAAAADDEEGGGG", "----------**--------------------MMMM----------**---M------------"},
  {3, "The Yeast Mitochondrial Code", "FFLLSSSSYY**CCWWTTTTPPPPHHQQRRRRIIMMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "----------**----------------------MM---------------M------------"},
  {4, "The Mold, Protozoan, and Coelenterate Mitochondrial Code and the Mycoplasma/Spiroplasma Code", "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "--MM------**-------M------------MMMM---------------M------------"},
  {5, "The Invertebrate Mitochondrial Code", "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSSSVVVVAAAADDEEGGGG", "---M------**--------------------MMMM---------------M------------"},
  {6, "The Ciliate, Dasycladacean and Hexamita Nuclear Code", "FFLLSSSSYYQQCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "--------------*--------------------M----------------------------"},
  {9, "The Echinoderm and Flatworm Mitochondrial Code", "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG", "----------**-----------------------M---------------M------------"},
  {10, "The Euplotid Nuclear Code", "FFLLSSSSYY**CCCWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "----------**-----------------------M----------------------------"},
  {11, "The Bacterial, Archaeal and Plant Plastid Code", "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "---M------**--*----M------------MMMM---------------M------------"},
  {12, "The Alternative Yeast Nuclear Code", "FFLLSSSSYY**CC*WLLLSPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "----------**--*----M---------------M----------------------------"},
  {13, "The Ascidian Mitochondrial Code", "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSGGVVVVAAAADDEEGGGG", "---M------**----------------------MM---------------M------------"},
  {14, "The Alternative Flatworm Mitochondrial Code", "FFLLSSSSYYY*CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG", "-----------*-----------------------M----------------------------"},
  {16, "Chlorophycean Mitochondrial Code", "FFLLSSSSYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "----------*---*--------------------M----------------------------"},
  {21, "Trematode Mitochondrial Code", "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNNKSSSSVVVVAAAADDEEGGGG", "----------**-----------------------M---------------M------------"},
  {22, "Scenedesmus obliquus Mitochondrial Code", "FFLLSS*SYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "------*---*---*--------------------M----------------------------"},
  {23, "Thraustochytrium Mitochondrial Code", "FF*LSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "--*-------**--*-----------------M--M---------------M------------"},
  {24, "Pterobranchia Mitochondrial Code", "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSSKVVVVAAAADDEEGGGG", "---M------**-------M---------------M---------------M------------"},
  {25, "Candidate Division SR1 and Gracilibacteria Code", "FFLLSSSSYY**CCGWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "---M------**-----------------------M---------------M------------"},
  {26, "Pachysolen tannophilus Nuclear Code", "FFLLSSSSYY**CC*WLLLAPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "----------**--*----M---------------M----------------------------"},
  {27, "Karyorelict Nuclear", "FFLLSSSSYYQQCCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "--------------*--------------------M----------------------------"},
  {28, "Condylostoma Nuclear", "FFLLSSSSYYQQCCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "----------**--*--------------------M----------------------------"},
  {29, "Mesodinium Nuclear", "FFLLSSSSYYYYCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "--------------*--------------------M----------------------------"},
  {30, "Peritrich Nuclear", "FFLLSSSSYYEECC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "--------------*--------------------M----------------------------"},
  {31, "Blastocrithidia Nuclear", "FFLLSSSSYYEECCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "----------**-----------------------M----------------------------"},
};
static const gcode_t *gcode_find(int id) {
  for (size_t i = 0; i < sizeof gcodes / sizeof gcodes[0]; i++) if (gcodes[i].id == id) return &gcodes[i];
  return NULL;
}
static int iupac_mask(uint8_t c) { /* bit0=T bit1=C bit2=A bit3=G (TCAG order) */
  switch (c | 32) {
    case 't': case 'u': return 1; case 'c': return 2; case 'a': return 4; case 'g': return 8;
    case 'r': return 4 | 8; case 'y': return 2 | 1; case 's': return 2 | 8; case 'w': return 4 | 1;
    case 'k': return 8 | 1; case 'm': return 4 | 2; case 'b': return 2 | 8 | 1; case 'd': return 4 | 8 | 1;
    case 'h': return 4 | 2 | 1; case 'v': return 4 | 2 | 8; case 'n': return 15;
  }
  return 0;
}
/* amino acid for a (possibly ambiguous) codon: all expansions agree -> aa, else 'X';
 * "---" -> '-'; any non-IUPAC byte -> unknown (-1).  is_init: every expansion is a start codon. */
static int codon_aa(const gcode_t *t, const uint8_t *c, int *is_init) {
  if (is_init) *is_init = 0;
  if (c[0] == '-' && c[1] == '-' && c[2] == '-') return '-';
  int m0 = iupac_mask(c[0]), m1 = iupac_mask(c[1]), m2 = iupac_mask(c[2]);
  if (!m0 || !m1 || !m2) return -1;
  int aa = 0, init = 1;
  for (int i = 0; i < 4; i++) if (m0 >> i & 1)
    for (int j = 0; j < 4; j++) if (m1 >> j & 1)
      for (int k = 0; k < 4; k++) if (m2 >> k & 1) {
        int idx = i * 16 + j * 4 + k;
        int a = t->aas[idx];
        if (t->starts[idx] != 'M') init = 0;
        if (aa == 0) aa = a; else if (aa != a) aa = 'X';
      }
  if (is_init) *is_init = init;
  return aa;
}
int orc_translate_codon(int table, const char *codon) {
  const gcode_t *t = gcode_find(table);
  if (!t || strlen(codon) != 3) return -1;
  return codon_aa(t, (const uint8_t *)codon, NULL);
}
static int parse_frames(const char *csv, int *frames, char *err) { /* lib/translate.go:46-61 */
  int nf = 0; const char *s = csv;
  while (*s) {
    char tok[32]; size_t k = 0;
    while (*s && *s != ',') { if (k < 31) tok[k++] = *s; s++; }
    tok[k] = 0; if (*s == ',') s++;
    char *endp; long f = strtol(tok, &endp, 10);
    if (k == 0 || *endp) { snprintf(err, 512, "invalid frame(s): %s. available: 1, 2, 3, -1, -2, -3, and 6 for all. multiple frames should be separated by comma", tok); return -1; }
    if (!(f == 1 || f == 2 || f == 3 || f == -1 || f == -2 || f == -3 || f == 6)) { snprintf(err, 512, "invalid frame: %ld. available: 1, 2, 3, -1, -2, -3, and 6 for all", f); return -1; }
    if (f == 6) { int all6[6] = {1, 2, 3, -1, -2, -3}; memcpy(frames, all6, sizeof all6); return 6; }
    if (nf < 4096) frames[nf++] = (int)f;  /* the reference takes any number of repeats; 4096 is this restatement's bound */
  }
  return nf;
}
int orc_translate(const uint8_t *data, size_t n, const orc_opts *o, orc_out *out) {
  memset(out, 0, sizeof *out);
  int ab = ab_from_type(o->SeqType, out->err);
  if (ab < 0) return -1;
  const gcode_t *t = gcode_find(o->TranslTable);
  if (!t) { snprintf(out->err, 512, "invalid translate table: %d", o->TranslTable); return -1; }
  int frames[4096];
  int nf = parse_frames(o->Frame ? o->Frame : "1", frames, out->err);
  if (nf < 0) return -1;
  parser_t p; parser_init(&p, data, n, ab, o);
  sink_t sink; memset(&sink, 0, sizeof sink);
  buf_t ob, rcb, prot; memset(&ob, 0, sizeof ob); memset(&rcb, 0, sizeof rcb); memset(&prot, 0, sizeof prot);
  int rc, once = 1;
  while ((rc = parser_read(&p)) > 0) {
    rec_t *r = &p.r;
    if (once) {
      int a = p.alphabet;
      if (!(a == AB_DNA || a == AB_DNARED || a == AB_RNA || a == AB_RNARED)) { snprintf(out->err, 512, "command 'seqkit translate' only apply to DNA/RNA sequences"); rc = -2; break; }
      once = 0;
    }
    for (int fi = 0; fi < nf && rc > 0; fi++) {
      int f = frames[fi];
      const uint8_t *s = r->seq; size_t l = r->seq_len;
      if (l < 3) { snprintf(out->err, 512, "seq: sequence too short to translate"); rc = -2; break; }
      if (f < 0) { revcom_into(&rcb, p.alphabet, r->seq, r->seq_len); s = rcb.p; }
      prot.n = 0;
      size_t start = (size_t)((f < 0 ? -f : f) - 1);
      for (size_t i = start; i + 3 <= l; i += 3) {
        int init, aa = codon_aa(t, s + i, &init);
        if (aa < 0) {
          if (o->AllowUnknownCodon) aa = 'X';
          else { snprintf(out->err, 512, "seq: unknown codon"); rc = -2; break; }
        }
        if (o->InitCodonAsM && i == start && init) aa = 'M';
        if (o->Clean && aa == '*') aa = 'X';
        buf_addc(&prot, (uint8_t)aa);
      }
      if (rc < 0) break;
      if (o->Trim) while (prot.n && (prot.p[prot.n - 1] == 'X' || prot.p[prot.n - 1] == '*')) prot.n--;
      ob.n = 0;
      buf_addc(&ob, '>');
      if (o->AppendFrame) { buf_add(&ob, r->id, r->id_len); buf_printf(&ob, "_frame=%d ", f); buf_add(&ob, r->desc, r->desc_len); } /* :134 */
      else buf_add(&ob, r->head, r->head_len);
      buf_addc(&ob, '\n');
      wrap_into(&ob, prot.p, prot.n, o->LineWidth);
      sink_elem(&sink, ob.p, ob.n);
    }
    if (rc < 0) break; /* lib/translate.go:126-131: `return nil, err` aborts the partition */
  }
  if (rc == -1) snprintf(out->err, 512, "%s", p.err);
  sink_to_out(&sink, out);
  free(ob.p); free(rcb.p); free(prot.p); parser_free(&p);
  return rc < 0 ? -1 : 0;
}

/* ------------------------------------------------------------------ locate
 * Locate.Before/Call exact default path (lib/locate.go:33-204, 395-769) with
 * rows in record order, then pattern order as given, '+' rows
 * then '-' rows; one '\n' per row; Protein/Unlimit input implies -P. */
static const uint8_t *find_sub(const uint8_t *h, size_t hn, const uint8_t *nd, size_t nn) {
  if (nn == 0) return h;
  if (hn < nn) return NULL;
  return (const uint8_t *)memmem(h, hn, nd, nn);
}
static int pattern_legal(const uint8_t *s, size_t n) { /* lib/locate.go:182-187, lib/grep.go:231-240 */
  return ab_is_valid(AB_DNARED, s, n) || ab_is_valid(AB_RNARED, s, n) || ab_is_valid(AB_PROTEIN, s, n);
}
static void locate_row(sink_t *sink, buf_t *ob, const orc_opts *o, const rec_t *r, const char *pname, const uint8_t *pat,
                       size_t plen, char strand, long begin, long end, const uint8_t *matched) {
  ob->n = 0;
  if (o->Gtf) { /* :617-627 */
    buf_add(ob, r->id, r->id_len);
    buf_printf(ob, "\tSeqKit\tlocation\t%ld\t%ld\t0\t%c\t.\tgene_id \"%s\"; ", begin, end, strand, pname);
  } else if (o->Bed) { /* :628-635 */
    buf_add(ob, r->id, r->id_len);
    buf_printf(ob, "\t%ld\t%ld\t%s\t0\t%c", begin - 1, end, pname, strand);
  } else { /* :637-654 */
    buf_add(ob, r->id, r->id_len);
    buf_printf(ob, "\t%s\t", pname); buf_add(ob, pat, plen);
    buf_printf(ob, "\t%c\t%ld\t%ld", strand, begin, end);
    if (!o->HideMatched) { buf_addc(ob, '\t'); buf_add(ob, matched, plen); }
  }
  sink_elem(sink, ob->p, ob->n);
}
int orc_locate(const uint8_t *data, size_t n, const orc_opts *o, orc_out *out) {
  memset(out, 0, sizeof *out);
  ab_init();
  int ab = ab_from_type(o->SeqType, out->err);
  if (ab < 0) return -1;
  if (o->n_patterns == 0) { snprintf(out->err, 512, "one of flags -p (--pattern) and -f (--pattern-file) needed"); return -1; }
  uint8_t **pats = (uint8_t **)calloc((size_t)o->n_patterns, sizeof(uint8_t *));
  size_t *plen = (size_t *)calloc((size_t)o->n_patterns, sizeof(size_t));
  int bad = 0;
  for (int i = 0; i < o->n_patterns; i++) {
    plen[i] = strlen(o->patterns[i]);
    pats[i] = (uint8_t *)malloc(plen[i] + 1); memcpy(pats[i], o->patterns[i], plen[i] + 1);
    if (o->IgnoreCase) lower_bytes(pats[i], plen[i]); /* :156-158 */
    if (plen[i] == 0) { snprintf(out->err, 512, "one of flags -p (--pattern) and -f (--pattern-file) needed"); bad = 1; break; }
    if (memchr(pats[i], '.', plen[i]) || !pattern_legal(pats[i], plen[i])) {
      snprintf(out->err, 512, "illegal DNA/RNA/Protein sequence: %s, you may switch on -d/--degenerate or -r/--use-regexp", o->pattern_names[i]);
      bad = 1; break;
    }
  }
  sink_t sink; memset(&sink, 0, sizeof sink);
  int rc = 0;
  if (!bad) {
    if (!(o->Gtf || o->Bed)) { /* :198-204 */
      const char *h = o->HideMatched ? "seqID\tpatternName\tpattern\tstrand\tstart\tend" : "seqID\tpatternName\tpattern\tstrand\tstart\tend\tmatched";
      sink_elem(&sink, (const uint8_t *)h, strlen(h));
    }
    parser_t p; parser_init(&p, data, n, ab, o);
    buf_t ob, fw, rv; memset(&ob, 0, sizeof ob); memset(&fw, 0, sizeof fw); memset(&rv, 0, sizeof rv);
    int only_pos = o->OnlyPositiveStrand, check = 1;
    while ((rc = parser_read(&p)) > 0) {
      rec_t *r = &p.r;
      if (check) { int a = parser_alphabet(&p); if (a == AB_UNLIMIT || a == AB_PROTEIN) only_pos = 1; check = 0; } /* :424-429 + Q8 */
      if (o->IgnoreCase) lower_bytes(r->seq, r->seq_len); /* :431-433 */
      long l = (long)r->seq_len;
      fw.n = 0; buf_add(&fw, r->seq, r->seq_len);
      if (o->Circular) buf_add(&fw, r->seq, r->seq_len); /* :437-439 */
      size_t sl = fw.n;
      for (int pi = 0; pi < o->n_patterns; pi++) {
        const uint8_t *pt = pats[pi]; long lp = (long)plen[pi];
        long offset = 0;
        for (;;) { /* :583-667 */
          if ((size_t)offset > sl) break;
          const uint8_t *hit = find_sub(fw.p + offset, sl - (size_t)offset, pt, (size_t)lp);
          if (!hit) break;
          long i = (long)(hit - (fw.p + offset));
          long begin = offset + i + 1;
          if (o->Circular && begin > l) break;
          long end = offset + i + lp;
          locate_row(&sink, &ob, o, r, o->pattern_names[pi], pt, (size_t)lp, '+', begin, end, fw.p + begin - 1);
          offset = o->NonGreedy ? offset + i + lp + 1 : offset + i + 1;
          if (offset >= (long)sl) break;
        }
        if (only_pos) continue;
        revcom_into(&rv, p.alphabet, fw.p, sl); /* :673, recomputed per pattern */
        offset = 0;
        for (;;) { /* :679-766 */
          if ((size_t)offset > sl) break;
          const uint8_t *hit = find_sub(rv.p + offset, sl - (size_t)offset, pt, (size_t)lp);
          if (!hit) break;
          long i = (long)(hit - (rv.p + offset));
          if (o->Circular && offset + i + 1 > l) break;
          long begin = l - offset - (i + lp) + 1, end = l - offset - i;
          if (offset + i + lp > l) { begin += l; end += l; }
          locate_row(&sink, &ob, o, r, o->pattern_names[pi], pt, (size_t)lp, '-', begin, end, rv.p + offset + i);
          offset = o->NonGreedy ? offset + i + lp + 1 : offset + i + 1;
          if (offset >= (long)sl) break;
        }
      }
    }
    if (rc < 0) snprintf(out->err, 512, "%s", p.err);
    free(ob.p); free(fw.p); free(rv.p); parser_free(&p);
  }
  sink_to_out(&sink, out);
  for (int i = 0; i < o->n_patterns; i++) free(pats[i]);
  free(pats); free(plen);
  return (bad || rc < 0) ? -1 : 0;
}

/* -------------------------------------------------------------------- grep
 * Grep.Before + grepGeneral, non-regexp zero-mismatch path (lib/grep.go:41-253,
 * 367-542).  --delete-matched / -r / -d / -m are out of scope. */
static int parse_region(const char *region, int *start, int *end, const char *cmd, char *err) { /* lib/grep.go:93-118, lib/subseq.go:78-96 */
  const char *c = region; int ok = 1;
  if (*c == '-') c++;
  if (!(*c >= '0' && *c <= '9')) ok = 0;
  while (*c >= '0' && *c <= '9') c++;
  if (*c != ':') ok = 0; else c++;
  if (*c == '-') c++;
  if (!(*c >= '0' && *c <= '9')) ok = 0;
  while (*c >= '0' && *c <= '9') c++;
  if (*c) ok = 0;
  if (!ok) { snprintf(err, 512, "invalid region: %s. type \"seqkit %s -h\" for more examples", region, cmd); return -1; }
  *start = atoi(region); *end = atoi(strchr(region, ':') + 1);
  if (*start == 0 || *end == 0) { snprintf(err, 512, "both start and end should not be 0"); return -1; }
  if (*start < 0 && *end > 0) { snprintf(err, 512, "when start < 0, end should not > 0"); return -1; }
  return 0;
}
int orc_grep(const uint8_t *data, size_t n, const orc_opts *o, orc_out *out) {
  memset(out, 0, sizeof *out);
  ab_init();
  int ab = ab_from_type(o->SeqType, out->err);
  if (ab < 0) return -1;
  if (o->n_patterns == 0) { snprintf(out->err, 512, "one of flags -p (--pattern) and -f (--pattern-file) needed"); return -1; }
  int by_seq = o->BySeq, limit_region = 0, rstart = 0, rend = 0;
  if (o->Region && o->Region[0]) { /* :93-118 */
    limit_region = 1; by_seq = 1;
    if (parse_region(o->Region, &rstart, &rend, "grep", out->err)) return -1;
  }
  int np = 0;
  uint8_t **pats = (uint8_t **)calloc((size_t)o->n_patterns, sizeof(uint8_t *));
  size_t *plen = (size_t *)calloc((size_t)o->n_patterns, sizeof(size_t));
  int bad = 0;
  for (int i = 0; i < o->n_patterns; i++) { /* :199-249 */
    size_t l = strlen(o->patterns[i]);
    if (l == 0) continue; /* pattern file: empty lines skipped (:136-138) */
    if (by_seq && !pattern_legal((const uint8_t *)o->patterns[i], l)) { snprintf(out->err, 512, "illegal DNA/RNA/Protein sequence: %s", o->patterns[i]); bad = 1; break; }
    pats[np] = (uint8_t *)malloc(l + 1); memcpy(pats[np], o->patterns[i], l + 1); plen[np] = l;
    if (o->IgnoreCase) lower_bytes(pats[np], l);
    np++;
  }
  sink_t sink; memset(&sink, 0, sizeof sink);
  int rc = 0;
  if (!bad) {
    parser_t p; parser_init(&p, data, n, ab, o);
    buf_t ob, tg, rv; memset(&ob, 0, sizeof ob); memset(&tg, 0, sizeof tg); memset(&rv, 0, sizeof rv);
    int only_pos = o->OnlyPositiveStrand, check = 1;
    uint64_t count = 0;
    while ((rc = parser_read(&p)) > 0) {
      rec_t *r = &p.r;
      if (check) { int a = parser_alphabet(&p); if (a == AB_UNLIMIT || a == AB_PROTEIN) only_pos = 1; check = 0; } /* :404-409 */
      int hit = 0;
      for (int strand = 0; strand < 2 && !hit; strand++) { /* :427-514 */
        if (strand == 1 && (!by_seq || only_pos)) break;
        const uint8_t *target; size_t tl;
        if (by_seq) {
          const uint8_t *s = r->seq; size_t l = r->seq_len;
          if (strand == 1) { revcom_into(&rv, p.alphabet, r->seq, r->seq_len); s = rv.p; }
          tg.n = 0;
          if (limit_region) { size_t s0, sl = orc_subseq_range(l, rstart, rend, &s0); buf_add(&tg, s + s0, sl); }
          else if (o->Circular) { buf_add(&tg, s, l); buf_add(&tg, s, l); }
          else buf_add(&tg, s, l);
          if (o->IgnoreCase) lower_bytes(tg.p, tg.n);
          target = tg.p; tl = tg.n;
          for (int k = 0; k < np; k++) if (find_sub(target, tl, pats[k], plen[k])) { hit = 1; break; } /* :474-482 */
        } else {
          const uint8_t *s = o->ByName ? r->head : r->id; size_t l = o->ByName ? r->head_len : r->id_len;
          tg.n = 0; buf_add(&tg, s, l);
          if (o->IgnoreCase) lower_bytes(tg.p, tg.n);
          for (int k = 0; k < np; k++) if (plen[k] == l && memcmp(pats[k], tg.p, l) == 0) { hit = 1; break; } /* :501-512 */
        }
      }
      if (o->InvertMatch ? hit : !hit) continue; /* :516-524 */
      if (o->Count) { count++; continue; }
      int fq = p.is_fastq;
      format_record(&ob, r->head, r->head_len, r->seq, r->seq_len, r->qual, r->qual_len, fq, fq ? 0 : o->LineWidth);
      sink_elem(&sink, ob.p, ob.n - 1); /* :529-533 */
    }
    if (rc < 0) snprintf(out->err, 512, "%s", p.err);
    else if (o->Count) { char t[32]; int k = snprintf(t, sizeof t, "%llu", (unsigned long long)count); sink_elem(&sink, (uint8_t *)t, (size_t)k); } /* :538-540 */
    free(ob.p); free(tg.p); free(rv.p); parser_free(&p);
  }
  sink_to_out(&sink, out);
  for (int i = 0; i < np; i++) free(pats[i]);
  free(pats); free(plen);
  return (bad || rc < 0) ? -1 : 0;
}

/* ------------------------------------------------------------------ subseq
 * SubseqTransform region mode (lib/subseq.go:36-96,167-190,314-317), Q8: one '\n' per element. */
int orc_subseq(const uint8_t *data, size_t n, const orc_opts *o, orc_out *out) {
  memset(out, 0, sizeof *out);
  int ab = ab_from_type(o->SeqType, out->err);
  if (ab < 0) return -1;
  int start, end;
  if (!o->Region || !o->Region[0]) { snprintf(out->err, 512, "one of the options needed: -r/--region, --bed, --gtf"); return -1; }
  if (parse_region(o->Region, &start, &end, "subseq", out->err)) return -1;
  parser_t p; parser_init(&p, data, n, ab, o);
  sink_t sink; memset(&sink, 0, sizeof sink);
  buf_t ob; memset(&ob, 0, sizeof ob);
  int rc;
  while ((rc = parser_read(&p)) > 0) {
    rec_t *r = &p.r;
    size_t s0, sl = orc_subseq_range(r->seq_len, start, end, &s0);
    int fq = p.is_fastq;
    format_record(&ob, r->head, r->head_len, r->seq + s0, sl, r->qual_len ? r->qual + s0 : r->qual, r->qual_len ? sl : 0, fq,
                  fq ? 0 : o->LineWidth);
    sink_elem(&sink, ob.p, ob.n - 1);
  }
  if (rc < 0) snprintf(out->err, 512, "%s", p.err);
  sink_to_out(&sink, out);
  free(ob.p); parser_free(&p);
  return rc < 0 ? -1 : 0;
}

/* ------------------------------------------------------------------- fq2fa
 * Fq2Fa.Call (lib/fq2fa.go:36-61): record.Seq.Qual = []byte{}; Format(0) minus the final '\n'. */
int orc_fq2fa(const uint8_t *data, size_t n, const orc_opts *o, orc_out *out) {
  memset(out, 0, sizeof *out);
  int ab = ab_from_type(o->SeqType, out->err);
  if (ab < 0) return -1;
  parser_t p; parser_init(&p, data, n, ab, o);
  sink_t sink; memset(&sink, 0, sizeof sink);
  buf_t ob; memset(&ob, 0, sizeof ob);
  int rc;
  while ((rc = parser_read(&p)) > 0) {
    rec_t *r = &p.r;
    format_record(&ob, r->head, r->head_len, r->seq, r->seq_len, r->qual, 0, 0, 0); /* :53-54 */
    sink_elem(&sink, ob.p, ob.n - 1);                                               /* :56 */
  }
  if (rc < 0) snprintf(out->err, 512, "%s", p.err);
  sink_to_out(&sink, out);
  free(ob.p); parser_free(&p);
  return rc < 0 ? -1 : 0;
}

/* ------------------------------------------------------- duplicate / range / head
 * Operators on the RAW elements (record text minus one trailing '\n'); FileStore writes element + '\n'.
 *   Duplicate.Call      lib/duplicate.go:24-30   every element `times` times
 *   RangePrepare.Call   lib/range.go:26-31       element kept iff start <= index < end, index = global 0-based position
 *   RangeFilter.Call    lib/range.go:42-44       (MapWithIndex; drops the "" placeholders)
 *   Head                bigseqkit/head.go:33-44   = Range("1:N"), i.e. start 0, end N
 * The snapshot's driver rejects every start <= end (bigseqkit/range.go:85), so range / head cannot run there; the
 * library operators are restated with the semantics their code states. */
static void raw_elem(sink_t *sink, const uint8_t *d, uint64_t a, uint64_t b) {
  if (b > a && d[b - 1] == '\n') b--;
  sink_elem(sink, d + a, b - a);
}
int orc_duplicate(const uint8_t *data, size_t n, int64_t times, orc_out *out) {
  memset(out, 0, sizeof *out);
  uint64_t *st; size_t cnt = orc_frame(data, n, &st);
  sink_t sink; memset(&sink, 0, sizeof sink);
  for (size_t r = 0; r < cnt; r++)
    for (int64_t t = 0; t < times; t++) raw_elem(&sink, data, st[r], st[r + 1]);
  sink_to_out(&sink, out);
  free(st);
  return 0;
}
int orc_range(const uint8_t *data, size_t n, int64_t start, int64_t end, int64_t index_base, orc_out *out) {
  memset(out, 0, sizeof *out);
  uint64_t *st; size_t cnt = orc_frame(data, n, &st);
  sink_t sink; memset(&sink, 0, sizeof sink);
  for (size_t r = 0; r < cnt; r++) {
    const int64_t idx = index_base + (int64_t)r;
    if (start <= idx && idx < end) raw_elem(&sink, data, st[r], st[r + 1]);
  }
  sink_to_out(&sink, out);
  free(st);
  return 0;
}

/* ------------------------------------------------- multi-threaded CPU baseline
 * Record-aligned byte-range shards, one thread each (the restatement of
 * "IgnisHPC partitions x executor threads").  rmdup: parallel Prepare, keys
 * partitioned by key % T (GroupByKey), parallel Check, parallel Format. */
typedef struct {
  const char *op; const uint8_t *d; size_t n; const orc_opts *o;
  uint64_t nrec, out_bytes; int rc;
  orc_stats st;
  int keep_out; orc_out out; /* keep_out: the shard's output is handed to the caller (orc_run_mt_out) */
} mt_job;
typedef int (*map_fn)(const uint8_t *, size_t, const orc_opts *, orc_out *);
static map_fn mt_map_fn(const char *op) {
  if (!strcmp(op, "seq")) return orc_seq;
  if (!strcmp(op, "translate")) return orc_translate;
  if (!strcmp(op, "locate")) return orc_locate;
  if (!strcmp(op, "grep")) return orc_grep;
  if (!strcmp(op, "subseq")) return orc_subseq;
  if (!strcmp(op, "fq2fa")) return orc_fq2fa;
  return NULL;
}
static void *mt_worker(void *arg) {
  mt_job *j = (mt_job *)arg;
  map_fn f = mt_map_fn(j->op);
  if (!strcmp(j->op, "stats")) {
    j->rc = orc_stats_run(j->d, j->n, j->o, &j->st);
    j->nrec = j->st.num; j->out_bytes = 0;
  } else if (f) {
    j->rc = f(j->d, j->n, j->o, &j->out);
    j->nrec = j->out.n_elem; j->out_bytes = j->out.n;
    if (!j->keep_out) orc_out_free(&j->out);
  } else j->rc = -1;
  return NULL;
}
/* rmdup multi-thread pieces */
typedef struct {
  const uint8_t *d; size_t n; const orc_opts *o;
  size_t rec0;             /* global index of first record of the shard */
  uint64_t *keys;          /* global arrays */
  const uint8_t **subj; uint32_t *subj_len; uint8_t *keep;
  size_t n_total; int tid, nthreads;
  uint64_t nrec, out_bytes; int rc;
  buf_t own;               /* lower-cased subject copies when IgnoreCase */
  int keep_out; buf_t all; uint64_t *eoff; /* keep_out: kept records (each + '\n') and their start offsets */
} dd_job;
static void *dd_prepare(void *arg) {
  dd_job *j = (dd_job *)arg;
  char err[512]; int ab = ab_from_type(j->o->SeqType, err);
  parser_t p; parser_init(&p, j->d, j->n, ab, j->o);
  /* subjects must stay addressable after parsing: by-seq subjects of single-line
   * records are views into the input; otherwise copy */
  size_t c = j->rec0; int rc; buf_t tmp; memset(&tmp, 0, sizeof tmp);
  size_t *offs = (size_t *)malloc((p.n_rec + 1) * sizeof(size_t));
  size_t k = 0;
  while ((rc = parser_read(&p)) > 0) {
    const uint8_t *s; size_t l;
    rmdup_subject(j->o, &p.r, &tmp, &s, &l);
    j->keys[c] = orc_xxh64(s, l, 0);
    offs[k++] = j->own.n; buf_add(&j->own, s, l);
    j->subj_len[c] = (uint32_t)l;
    c++;
  }
  for (size_t i = 0; i < k; i++) j->subj[j->rec0 + i] = j->own.p + offs[i];
  free(offs); free(tmp.p); parser_free(&p);
  j->rc = rc < 0 ? -1 : 0;
  return NULL;
}
static void *dd_check(void *arg) {
  dd_job *j = (dd_job *)arg;
  size_t mine = 0;
  for (size_t i = 0; i < j->n_total; i++) if (j->keys[i] % (uint64_t)j->nthreads == (uint64_t)j->tid) mine++;
  size_t cap = 64; while (cap < 2 * mine + 2) cap *= 2;
  uint32_t *tab = (uint32_t *)calloc(cap, sizeof(uint32_t)); /* record index + 1 */
  for (size_t i = 0; i < j->n_total; i++) {
    uint64_t key = j->keys[i];
    if (key % (uint64_t)j->nthreads != (uint64_t)j->tid) continue;
    size_t h = (size_t)((key / (uint64_t)j->nthreads) & (cap - 1));
    int dup = 0;
    for (;;) {
      if (!tab[h]) break;
      size_t q = tab[h] - 1;
      if (j->keys[q] == key && j->subj_len[q] == j->subj_len[i] && memcmp(j->subj[q], j->subj[i], j->subj_len[i]) == 0) { dup = 1; break; }
      h = (h + 1) & (cap - 1);
    }
    if (dup) j->keep[i] = 0; else { tab[h] = (uint32_t)(i + 1); j->keep[i] = 1; }
  }
  free(tab);
  return NULL;
}
static void *dd_format(void *arg) {
  dd_job *j = (dd_job *)arg;
  char err[512]; int ab = ab_from_type(j->o->SeqType, err);
  parser_t p; parser_init(&p, j->d, j->n, ab, j->o);
  buf_t ob, all; memset(&ob, 0, sizeof ob); memset(&all, 0, sizeof all);
  size_t c = j->rec0; int rc;
  if (j->keep_out) j->eoff = (uint64_t *)malloc((p.n_rec + 1) * sizeof(uint64_t));
  while ((rc = parser_read(&p)) > 0) {
    if (j->keep[c++]) {
      rec_t *r = &p.r; int fq = p.is_fastq;
      format_record(&ob, r->head, r->head_len, r->seq, r->seq_len, r->qual, r->qual_len, fq, fq ? 0 : j->o->LineWidth);
      if (j->keep_out) j->eoff[j->nrec] = all.n;
      buf_add(&all, ob.p, ob.n);
      j->nrec++;
    }
  }
  j->out_bytes = all.n;
  free(ob.p); parser_free(&p);
  if (j->keep_out) j->all = all; else free(all.p);
  return NULL;
}
static int run_mt_impl(const char *op, const uint8_t *data, size_t n, const orc_opts *o, int threads, uint64_t *n_records, uint64_t *out_bytes,
                       orc_out *out, orc_stats *st_out) {
  if (out) memset(out, 0, sizeof *out);
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  /* record-aligned cuts: move each cut forward to the next record start */
  size_t cut[257]; int fq = n > 0 && data[0] == '@';
  cut[0] = 0;
  for (int t = 1; t < threads; t++) {
    size_t c = n / (size_t)threads * (size_t)t;
    if (c < cut[t - 1]) c = cut[t - 1];
    while (c + 1 < n) {
      if (data[c] == '\n') {
        int hit = fq ? (data[c + 1] == '@' && !(c >= 2 && data[c - 2] == '\n' && data[c - 1] == '+')) : data[c + 1] == '>';
        if (hit) break;
      }
      c++;
    }
    cut[t] = c + 1 < n ? c + 1 : n;
  }
  cut[threads] = n;
  pthread_t th[256];
  uint64_t nr = 0, ob = 0; int rc = 0;
  if (!strcmp(op, "rmdup")) {
    dd_job *jobs = (dd_job *)calloc((size_t)threads, sizeof(dd_job));
    size_t total = 0;
    for (int t = 0; t < threads; t++) {
      uint64_t *st; size_t c = orc_frame(data + cut[t], cut[t + 1] - cut[t], &st); free(st);
      jobs[t].rec0 = total; total += c;
    }
    uint64_t *keys = (uint64_t *)malloc((total + 1) * 8);
    const uint8_t **subj = (const uint8_t **)malloc((total + 1) * sizeof(void *));
    uint32_t *sl = (uint32_t *)malloc((total + 1) * 4);
    uint8_t *keep = (uint8_t *)calloc(total + 1, 1);
    for (int t = 0; t < threads; t++) {
      jobs[t].d = data + cut[t]; jobs[t].n = cut[t + 1] - cut[t]; jobs[t].o = o; jobs[t].keys = keys; jobs[t].subj = subj;
      jobs[t].subj_len = sl; jobs[t].keep = keep; jobs[t].n_total = total; jobs[t].tid = t; jobs[t].nthreads = threads;
      jobs[t].keep_out = out != NULL;
    }
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, dd_prepare, &jobs[t]);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, dd_check, &jobs[t]);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, dd_format, &jobs[t]);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    for (int t = 0; t < threads; t++) { nr += jobs[t].nrec; ob += jobs[t].out_bytes; if (jobs[t].rc) rc = -1; free(jobs[t].own.p); }
    if (out) { /* kept records of all shards in input order */
      out->data = (uint8_t *)malloc(ob + 1); out->elem_off = (uint64_t *)malloc((nr + 1) * sizeof(uint64_t));
      size_t pos = 0, e = 0;
      for (int t = 0; t < threads; t++) {
        for (size_t i = 0; i < jobs[t].nrec; i++) out->elem_off[e++] = pos + jobs[t].eoff[i];
        if (jobs[t].all.n) memcpy(out->data + pos, jobs[t].all.p, jobs[t].all.n);
        pos += jobs[t].all.n; free(jobs[t].all.p); free(jobs[t].eoff);
      }
      out->elem_off[e] = pos; out->n = pos; out->n_elem = e;
    }
    nr = total;
    free(keys); free(subj); free(sl); free(keep); free(jobs);
  } else {
    mt_job *jobs = (mt_job *)calloc((size_t)threads, sizeof(mt_job));
    if (strcmp(op, "stats") && !mt_map_fn(op)) { free(jobs); return -1; }
    for (int t = 0; t < threads; t++) { jobs[t].op = op; jobs[t].d = data + cut[t]; jobs[t].n = cut[t + 1] - cut[t]; jobs[t].o = o; jobs[t].keep_out = out != NULL; }
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    orc_stats merged; memset(&merged, 0, sizeof merged);
    for (int t = 0; t < threads; t++) {
      nr += jobs[t].nrec; ob += jobs[t].out_bytes; if (jobs[t].rc) rc = -1;
      if (!strcmp(op, "stats")) { orc_stats_merge(&merged, &jobs[t].st); orc_stats_free(&jobs[t].st); }
    }
    if (!strcmp(op, "stats")) {
      orc_stats_finalise(&merged, o->All); nr = merged.num; ob = merged.sum_len;
      if (st_out) *st_out = merged; else orc_stats_free(&merged);
    } else if (out) {
      /* elements of all shards in input order; locate prints its header row in partition 0 only (lib/locate.go:198-204) */
      const int hdr = !strcmp(op, "locate") && !(o->Gtf || o->Bed);
      size_t tot = 0, ne = 0;
      for (int t = 0; t < threads; t++) { tot += jobs[t].out.n; ne += jobs[t].out.n_elem; }
      out->data = (uint8_t *)malloc(tot + 1); out->elem_off = (uint64_t *)malloc((ne + 1) * sizeof(uint64_t));
      size_t pos = 0, e = 0;
      for (int t = 0; t < threads; t++) {
        orc_out *so = &jobs[t].out;
        size_t e0 = (hdr && t > 0 && so->n_elem) ? 1 : 0;
        if (jobs[t].rc && !out->err[0]) memcpy(out->err, so->err, sizeof out->err);
        if (so->n_elem > e0) {
          const uint64_t b0 = so->elem_off[e0];
          for (size_t i = e0; i < so->n_elem; i++) out->elem_off[e++] = pos + (so->elem_off[i] - b0);
          memcpy(out->data + pos, so->data + b0, so->n - b0);
          pos += so->n - b0;
        }
        orc_out_free(so);
      }
      out->elem_off[e] = pos; out->n = pos; out->n_elem = e; ob = pos;
    }
    free(jobs);
  }
  if (n_records) *n_records = nr;
  if (out_bytes) *out_bytes = ob;
  return rc;
}
int orc_run_mt(const char *op, const uint8_t *data, size_t n, const orc_opts *o, int threads, uint64_t *n_records, uint64_t *out_bytes) {
  return run_mt_impl(op, data, n, o, threads, n_records, out_bytes, NULL, NULL);
}
/* as orc_run_mt, but the output of the whole input is returned (shards concatenated in input order): the full-size
 * parity checks of bench.py / tests compare it with the CUDA library's output byte for byte.  op "stats": *st receives
 * the merged, finalised result (out stays empty).  n_records = input records (stats, rmdup) or elements (map operators). */
int orc_run_mt_out(const char *op, const uint8_t *data, size_t n, const orc_opts *o, int threads, orc_out *out, orc_stats *st,
                   uint64_t *n_records) {
  uint64_t ob = 0;
  orc_stats tmp; memset(&tmp, 0, sizeof tmp);
  int rc = run_mt_impl(op, data, n, o, threads, n_records, &ob, out, st ? st : &tmp);
  if (!st) orc_stats_free(&tmp);
  return rc;
}
