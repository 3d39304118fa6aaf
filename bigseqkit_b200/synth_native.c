/*
 * synth_native.c -- fast deterministic generators of the BASELINE.json synthetic inputs, for the
 * benches and the full-size tests (numpy takes 20-30 s per GiB; this takes well under a second per GiB on a few
 * threads).  Counter-based: every field of record i is a pure function of (seed, i), so threads fill disjoint
 * record ranges of the caller's buffer directly (e.g. a pinned staging buffer) and the result does not depend on
 * the thread count.  Not part of the hot path; plain C, no CUDA.
 *
 *   bsk_synth_fastq    C2 / C3  4-line FASTQ, "@SIM:1:FC:<lane>:<tile>:<x>:<y> 1:N:0:ACGT", read_len bases over
 *                               ACGT with ~0.1 % N, Phred+33 in [2, 40]; with dup_ppm > 0 record i copies the
 *                               SEQUENCE of a uniformly chosen earlier record with that probability (chains
 *                               resolve to the original, so a copy of a copy equals the original)
 *   bsk_synth_fasta_reads  C1   ">seq%08d" + read_len bases on one line
 *   bsk_synth_contigs  C4       ">contig%06d len=%d", log-uniform lengths, wrapped at `width`
 *   bsk_synth_cds      C5       ">cds%07d gene=g%d", lengths uniform multiples of 3, starts with ATG, wrapped
 *
 * Every function returns the number of bytes written (whole records only, <= cap) and stores the record count.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t mix64(uint64_t x) { /* splitmix64 finaliser */
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
static inline uint64_t rnd(uint64_t seed, uint64_t field, uint64_t i, uint64_t j) {
  return mix64(mix64(seed * 0xD6E8FEB86659FD93ull + field) ^ mix64(i * 0xA24BAED4963EE407ull + j));
}
static inline uint64_t below(uint64_t r, uint64_t n) { return (uint64_t)(((__uint128_t)r * n) >> 64); }

static int put_uint(uint8_t *d, uint64_t v, int width) { /* decimal, zero padded to width (0 = none) */
  char t[24];
  int n = 0;
  do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n < width) t[n++] = '0';
  for (int k = 0; k < n; k++) d[k] = (uint8_t)t[n - 1 - k];
  return n;
}
static void put_bases(uint8_t *d, uint64_t seed, uint64_t rec, uint64_t first, uint64_t len) { /* bases [first, first+len) */
  static const char acgt[4] = {'A', 'C', 'G', 'T'};
  uint64_t j = first, end = first + len;
  while (j < end) {
    uint64_t w = rnd(seed, 11, rec, j >> 5);
    unsigned k = (unsigned)(j & 31);
    w >>= 2 * k;
    for (; k < 32 && j < end; k++, j++, w >>= 2) *d++ = (uint8_t)acgt[w & 3];
  }
}

typedef struct gen gen_t;
struct gen {
  uint64_t seed;
  int kind;             /* 0 fastq, 1 fasta reads, 2 contigs, 3 cds */
  uint32_t read_len, dup_ppm, width;
  uint32_t min_len, max_len;
  const uint64_t *off;  /* record offsets (n_rec + 1) */
  uint8_t *out;
};

/* ---- FASTQ */
static void fq_fields(const gen_t *g, uint64_t i, uint32_t f[4]) {
  const uint64_t r = rnd(g->seed, 1, i, 0), r2 = rnd(g->seed, 2, i, 0);
  f[0] = 1 + (uint32_t)below(r, 8);
  f[1] = 1101 + (uint32_t)below(mix64(r), 1128);
  f[2] = 1 + (uint32_t)below(r2, 29999);
  f[3] = 1 + (uint32_t)below(mix64(r2), 99998);
}
static int ndigits(uint32_t v) { int n = 1; while (v >= 10) { v /= 10; n++; } return n; }
static uint64_t fq_size(const gen_t *g, uint64_t i) {
  uint32_t f[4];
  fq_fields(g, i, f);
  return 10 + ndigits(f[0]) + 1 + ndigits(f[1]) + 1 + ndigits(f[2]) + 1 + ndigits(f[3]) + 12 + g->read_len + 3 + g->read_len + 1;
}
static uint64_t fq_root(const gen_t *g, uint64_t i) { /* record whose sequence record i carries */
  while (i > 0 && g->dup_ppm) {
    const uint64_t r = rnd(g->seed, 3, i, 0);
    if (below(r, 1000000) >= g->dup_ppm) break;
    i = below(rnd(g->seed, 4, i, 0), i);
  }
  return i;
}
static void fq_gen(const gen_t *g, uint64_t i, uint8_t *d) {
  uint32_t f[4];
  fq_fields(g, i, f);
  memcpy(d, "@SIM:1:FC:", 10); d += 10;
  for (int k = 0; k < 4; k++) { d += put_uint(d, f[k], 0); if (k < 3) *d++ = ':'; }
  memcpy(d, " 1:N:0:ACGT\n", 12); d += 12;
  const uint64_t root = fq_root(g, i);
  const uint32_t L = g->read_len;
  put_bases(d, g->seed, root, 0, L);
  { /* ~0.1 % N: a read carries one N with probability L / 1000 */
    const uint64_t r = rnd(g->seed, 5, root, 0);
    if (L && below(r, 1000) < L) d[below(r << 20 | r >> 44, L)] = 'N';
  }
  d += L;
  memcpy(d, "\n+\n", 3); d += 3;
  for (uint32_t j = 0; j < L; j += 8) {
    uint64_t w = rnd(g->seed, 6, i, j >> 3);
    for (uint32_t k = 0; k < 8 && j + k < L; k++, w >>= 8) d[j + k] = (uint8_t)(35 + (((uint32_t)(w & 0xff) * 39) >> 8));
  }
  d[L] = '\n';
}

/* ---- FASTA reads (fixed-size records) */
static uint64_t fa_size(const gen_t *g, uint64_t i) { (void)i; return 4 + 8 + 1 + (uint64_t)g->read_len + 1; }
static void fa_gen(const gen_t *g, uint64_t i, uint8_t *d) {
  memcpy(d, ">seq", 4); d += 4;
  d += put_uint(d, i % 100000000ull, 8);
  *d++ = '\n';
  put_bases(d, g->seed, i, 0, g->read_len);
  d[g->read_len] = '\n';
}

/* ---- wrapped FASTA (contigs, cds) */
static uint64_t wr_len(const gen_t *g, uint64_t i) {
  const uint64_t r = rnd(g->seed, 7, i, 0);
  if (g->kind == 2) { /* log-uniform in [min_len, max_len] */
    const double u = (double)(r >> 11) * (1.0 / 9007199254740992.0);
    const double l = exp(log((double)g->min_len) + u * (log((double)g->max_len) - log((double)g->min_len)));
    uint64_t L = (uint64_t)l;
    if (L < g->min_len) L = g->min_len;
    if (L > g->max_len) L = g->max_len;
    return L;
  }
  return 3 * (g->min_len / 3 + below(r, g->max_len / 3 - g->min_len / 3 + 1));
}
static int wr_header(const gen_t *g, uint64_t i, uint64_t L, uint8_t *d) {
  uint8_t *p = d;
  if (g->kind == 2) {
    memcpy(p, ">contig", 7); p += 7;
    p += put_uint(p, i, 6);
    memcpy(p, " len=", 5); p += 5;
    p += put_uint(p, L, 0);
  } else {
    memcpy(p, ">cds", 4); p += 4;
    p += put_uint(p, i, 7);
    memcpy(p, " gene=g", 7); p += 7;
    p += put_uint(p, i, 0);
  }
  *p++ = '\n';
  return (int)(p - d);
}
static uint64_t wr_size(const gen_t *g, uint64_t i) {
  uint8_t h[64];
  const uint64_t L = wr_len(g, i);
  return (uint64_t)wr_header(g, i, L, h) + L + (L + g->width - 1) / g->width;
}
static void wr_gen(const gen_t *g, uint64_t i, uint8_t *d) {
  const uint64_t L = wr_len(g, i);
  d += wr_header(g, i, L, d);
  for (uint64_t b = 0; b < L; b += g->width) {
    const uint64_t k = L - b < g->width ? L - b : g->width;
    put_bases(d, g->seed, i, b, k);
    if (b == 0 && g->kind == 3 && k >= 3) memcpy(d, "ATG", 3);
    d += k;
    *d++ = '\n';
  }
}

static uint64_t rec_size(const gen_t *g, uint64_t i) { return g->kind == 0 ? fq_size(g, i) : g->kind == 1 ? fa_size(g, i) : wr_size(g, i); }
static void rec_gen(const gen_t *g, uint64_t i, uint8_t *d) {
  if (g->kind == 0) fq_gen(g, i, d);
  else if (g->kind == 1) fa_gen(g, i, d);
  else wr_gen(g, i, d);
}

typedef struct { const gen_t *g; uint64_t r0, r1; } job_t;
static void *worker(void *arg) {
  const job_t *j = (const job_t *)arg;
  for (uint64_t i = j->r0; i < j->r1; i++) rec_gen(j->g, i, j->g->out + j->g->off[i]);
  return NULL;
}

static uint64_t generate(gen_t *g, uint8_t *out, uint64_t cap, uint64_t max_rec, int threads, uint64_t *n_rec_out) {
  uint64_t n_alloc = 1 << 16, n = 0, pos = 0;
  uint64_t *off = (uint64_t *)malloc((n_alloc + 1) * sizeof(uint64_t));
  for (;;) {
    if (max_rec && n >= max_rec) break;
    const uint64_t s = rec_size(g, n);
    if (pos + s > cap) break;
    if (n == n_alloc) { n_alloc *= 2; off = (uint64_t *)realloc(off, (n_alloc + 1) * sizeof(uint64_t)); }
    off[n++] = pos;
    pos += s;
  }
  off[n] = pos;
  g->off = off;
  g->out = out;
  if (threads < 1) threads = 1;
  if (threads > 64) threads = 64;
  pthread_t th[64];
  job_t jobs[64];
  /* split by bytes so that long contigs do not unbalance the threads */
  uint64_t r = 0;
  int nt = 0;
  for (int t = 0; t < threads && r < n; t++) {
    const uint64_t target = pos / (uint64_t)threads * (uint64_t)(t + 1);
    uint64_t e = r;
    while (e < n && (off[e + 1] <= target || e == r)) e++;
    if (t == threads - 1) e = n;
    jobs[nt].g = g; jobs[nt].r0 = r; jobs[nt].r1 = e;
    r = e;
    nt++;
  }
  for (int t = 0; t < nt; t++) pthread_create(&th[t], NULL, worker, &jobs[t]);
  for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
  free(off);
  if (n_rec_out) *n_rec_out = n;
  return pos;
}

uint64_t bsk_synth_fastq(uint8_t *out, uint64_t cap, uint64_t seed, uint32_t read_len, uint32_t dup_ppm, int threads, uint64_t *n_rec) {
  gen_t g; memset(&g, 0, sizeof g);
  g.seed = seed; g.kind = 0; g.read_len = read_len; g.dup_ppm = dup_ppm;
  return generate(&g, out, cap, 0, threads, n_rec);
}
uint64_t bsk_synth_fasta_reads(uint8_t *out, uint64_t cap, uint64_t seed, uint32_t read_len, uint64_t max_rec, int threads, uint64_t *n_rec) {
  gen_t g; memset(&g, 0, sizeof g);
  g.seed = seed; g.kind = 1; g.read_len = read_len;
  return generate(&g, out, cap, max_rec, threads, n_rec);
}
uint64_t bsk_synth_contigs(uint8_t *out, uint64_t cap, uint64_t seed, uint32_t min_len, uint32_t max_len, uint32_t width, int threads,
                           uint64_t *n_rec) {
  gen_t g; memset(&g, 0, sizeof g);
  g.seed = seed; g.kind = 2; g.min_len = min_len; g.max_len = max_len; g.width = width ? width : 60;
  return generate(&g, out, cap, 0, threads, n_rec);
}
uint64_t bsk_synth_cds(uint8_t *out, uint64_t cap, uint64_t seed, uint32_t min_len, uint32_t max_len, uint32_t width, int threads,
                       uint64_t *n_rec) {
  gen_t g; memset(&g, 0, sizeof g);
  g.seed = seed; g.kind = 3; g.min_len = min_len; g.max_len = max_len; g.width = width ? width : 60;
  return generate(&g, out, cap, 0, threads, n_rec);
}
