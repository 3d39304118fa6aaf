// kernels.h -- device data layout and host-side launchers of the CUDA kernels.
//
// HBM layout of one partition block (n < 4 GiB, so every offset is a u32):
//   in[n]            raw FASTA/FASTQ bytes as read from the file
//   ls[n_nl + 2]     line starts: ls[0] = 0, ls[k+1] = (k-th '\n') + 1, ls[n_nl+1] = n + 1
//                    -> line k holds bytes [ls[k], ls[k+1] - 1)
//   rl[n_rec + 1]    first line of every record, rl[n_rec] = number of real lines
//   per record (SoA, u32): head_off/len, seq_off/len (+ line run), qual_off/len (+ line run)
//   seq / qual arenas  only when some record spreads its sequence over several lines
//   out_len[n_rec+1] -> out_off[n_rec+1] (u64, exclusive scan) -> out[total]
#pragma once
#include <cstddef>
#include <cstdint>

#include "cuda_compat.h"

namespace bsk {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

// error kinds, ordered by the position of the check inside SeqParser.Read / Call
enum ErrKind : u32 { EK_VALIDATE = 0, EK_UNMATCHED = 1, EK_TOO_SHORT = 2, EK_UNKNOWN_CODON = 3 };
static const u64 kNoErr = ~0ull;

// device-resident status words of one call (mirrored to pinned host memory)
struct DevStatus {
  u64 err;         // min over (record << 8 | kind << 4 ...) ; kNoErr = none
  u64 aux;         // error detail (e.g. offending byte, lengths)
  u32 multiline;   // some record's sequence/quality spans several non-empty lines
  u32 guess_mask;  // AND of alphabet class masks over the guessed prefix of record 0
  u32 guess_len;   // number of bytes looked at
  u32 n_sel;       // scratch counter (selected items)
  u64 counters[8]; // op-specific totals (q20, q30, gaps, grep count, ...)
};

struct RecIndex {
  const u8 *in;
  u32 n;
  const u32 *ls;
  const u32 *rl;
  u32 n_lines;  // real lines (rl[n_rec])
  u32 n_rec;
  int fastq;
};

// per-record arrays produced by the parse kernel
struct RecArrays {
  u32 *head_off, *head_len;
  u32 *seq_line0, *seq_line1, *seq_off, *seq_len;      // line run [line0, line1) + contiguous offset
  u32 *qual_line0, *qual_line1, *qual_off, *qual_len;
};

// contiguous views consumed by the operator kernels
struct RecViews {
  const u8 *in, *seqb, *qualb;
  const u32 *name_off, *name_len;  // into in   (head, or ID when --only-id)
  const u32 *seq_off, *seq_len;    // into seqb
  const u32 *qual_off, *qual_len;  // into qualb
  u32 n_rec;
};

struct EmitCfg {
  u8 marker;  // '>' / '@' / 0
  u8 print_name, print_seq, print_qual, plus_line, reverse;
  u32 width;  // wrap width of the sequence lines (0 = single line)
};

namespace k {
// ---- record index (k_index.cu)
static const u32 kIndexTile = 16384;
void index_count(const u8 *in, u32 n, u64 *tile_cnt, u32 n_tiles, cudaStream_t s);
void index_fill(const u8 *in, u32 n, const u64 *tile_base, u32 *ls, u32 *rl, u32 n_tiles, cudaStream_t s);
void index_finish(u32 *ls, u32 *rl, u32 n, u32 n_nl, u32 n_rec, u32 n_lines, cudaStream_t s);
void parse_records(RecIndex ix, RecArrays ra, DevStatus *st, cudaStream_t s);
void squeeze_lines(RecIndex ix, RecArrays ra, const u32 *seq_aoff, const u32 *qual_aoff, u8 *seq_arena, u8 *qual_arena,
                   cudaStream_t s);
// uniformly wrapped FASTA: count the lines that break "all lines of a record as wide as its first, the last <= that",
// then squeeze arithmetically (byte q of a record's sequence = input byte seq_start + q + q / W)
void lines_uniform(RecIndex ix, u64 *n_bad, cudaStream_t s);
void squeeze_uniform(RecIndex ix, RecArrays ra, const u32 *seq_aoff, u8 *seq_arena, u32 total, cudaStream_t s);
void guess_alphabet(RecViews v, const u8 *class_mask, u32 limit, DevStatus *st, cudaStream_t s);
void id_desc(RecViews v, int id_ncbi, u32 *id_off, u32 *id_len, u32 *desc_off, u32 *desc_len, cudaStream_t s);
void validate_seq(RecViews v, const u8 *valid, u32 limit, DevStatus *st, cudaStream_t s);

// ---- generic record emitter (k_emit.cu)
void out_len(RecViews v, EmitCfg c, const u8 *keep, u32 *out_len, cudaStream_t s);
// in_limit / seq_limit / qual_limit: readable bytes of v.in / v.seqb / v.qualb (16-byte windows stop there)
void emit(RecViews v, EmitCfg c, const u64 *out_off, u8 *out, u64 total, const u8 *lut, u64 in_limit, u64 seq_limit,
          u64 qual_limit, cudaStream_t s);
// records whose formatted text is their input text: count the kept ones that are not, then compact byte ranges
void contig_check(RecViews v, EmitCfg c, const u8 *keep, int fastq, u32 in_bytes, u64 *n_bad, cudaStream_t s);
void emit_contig(RecViews v, const u64 *out_off, u8 *out, u64 total, u32 in_bytes, cudaStream_t s);

// ---- seq (k_seq.cu)
void remove_gaps(RecViews v, const u8 *gap, u8 *seq_out, u8 *qual_out, u32 *new_len, int has_qual, cudaStream_t s);
void seq_filter(RecViews v, int min_len, int max_len, double min_qual, double max_qual, const double *qual_pow,
                u8 *keep, cudaStream_t s);

// ---- FASTQ same-layout in-place kernel (k_fastq_inplace.cu): TMA-staged, no inter-CTA dependency
u32 fastq_inplace_tiles(u32 n);
u32 fastq_inplace_slot_stride();
void fastq_inplace(const u8 *in, u32 n, u8 *out, const u8 *lut, u32 *tile_cnt, u16 *slots, DevStatus *st, int reverse,
                   int use_lut, int group, u32 max_seg, u32 scan_halo, int n_sm, cudaStream_t s);
void fastq_elem_expand(const u32 *tile_cnt, const u64 *tile_base, const u16 *slots, u64 *elem_off, u32 n_tiles, u64 cap,
                       const DevStatus *st, cudaStream_t s);

// ---- stats on short records in one streaming pass (k_stats_tile.cu)
u32 stats_tile_bins();
void stats_tile(const u8 *in, u32 n, u64 *hist, DevStatus *st, int fastq, int all, int fq_offset, const u8 *gap_letters,
                int n_gap, u32 scan_halo, int n_sm, cudaStream_t s);

// ---- rmdup front half on 4-line FASTQ reads in one streaming pass (k_rmdup_tile.cu)
u32 rmdup_tile_tiles(u32 n);
u32 rmdup_tile_slot_stride();
size_t rmdup_tile_slot_bytes();
void rmdup_tile(const u8 *in, u32 n, void *slots, u32 *tile_cnt, DevStatus *st, int subject, int n_sm, cudaStream_t s);
void rmdup_tile_compact(const void *slots, const u32 *tile_cnt, const u64 *tile_base, u32 n_tiles, u64 *keys, u64 *fps,
                        RecArrays ra, u32 *id_len, cudaStream_t s);

// ---- locate with an equal-length ACGT panel on FASTA, one streaming pass over the raw bytes (k_locate_tile.cu)
struct LocateTileArgs {
  const u8 *in;
  u32 n, n_tiles;
  const u8 *lut;        // 256 byte classes (lt::C_*)
  const u32 *filter;    // first-level bitmap of the needle codes, locate_tile_filter_bits() bits, bit 31 - (i & 31) of word i >> 5
  const unsigned short *fptab;  // fingerprint table, 2^locate_tile_fp_bits() entries (copied to shared memory)
  u32 kmul, kmul2;      // bitmap index = (code * kmul) >> (32 - bits), fingerprint hash = code * kmul2; kmul_i = odd << (32 - 2L)
  u32 L, cmask;         // pattern length, mask of the 2L code bits
  u32 vmask, vbase;     // letters of the panel's case: (byte & vmask) must be one of vbase + {0, 2, 6, 0x13} in every byte
  const u32 *table;     // exact table: pairs {code, first needle + 1 (0 = empty)}, open addressing
  u32 tmask, tshift;
  const u32 *nd_code, *nd_ps;  // needles sorted by code: code, pattern << 1 | strand
  u32 n_needles;
  u64 *hitA, *hitB;     // raw hits: A = newlines of the tile in front << 32 | position of the last base, B = tile << 32 | ps
  u64 hit_cap;
  u64 *hdr_off, *hdr_nl;  // header lines: position of '>', newlines of its tile in front of it
  u64 hdr_cap;
  u32 *tile_nl;         // newlines per tile (owned range)
  DevStatus *st;        // counters[0] declined tiles, [5] hits, [6] header lines
};
u32 locate_tile_tiles(u32 n);
u32 locate_tile_bytes();
u32 locate_tile_filter_bits();
u32 locate_tile_fp_bits();
void locate_tile(LocateTileArgs a, int n_sm, cudaStream_t s);
void locate_records(const u8 *in, u32 n, const u64 *hdr_off, const u64 *hdr_nl, const u32 *tile_nl_base, u32 n_tiles, u32 n_rec,
                    u32 *name_off, u32 *name_len, u32 *seq_start, u32 *seq_nl, u32 *seq_len, cudaStream_t s);
void locate_resolve(u64 *hitA, u64 *hitB, u64 n_hits, const u64 *hdr_off, u32 n_rec, const u32 *tile_nl_base, const u32 *seq_start,
                    const u32 *seq_nl, const u32 *seq_len, u32 L, cudaStream_t s);

// ---- FASTA record table in one streaming pass over the raw bytes (k_fasta_tile.cu)
struct FastaTileArgs {
  const u8 *in;
  u32 n, n_tiles;
  u32 width;            // line width every record must be wrapped at (0: one sequence line per record)
  u8 *clean;            // one byte per 96-byte span: bit j = 16-byte chunk j holds only A, C, G, T, '\n'
  u64 *hdr_off, *hdr_nl;  // header lines: position of '>', newlines of its tile in front of it
  u64 hdr_cap;
  u32 *tile_nl;         // newlines per tile
  DevStatus *st;        // counters[0] tiles that break the line shape, [6] header lines, [7] no final newline
};
u32 fasta_tile_tiles(u32 n);
u32 fasta_tile_bytes();
u32 fasta_tile_spans_per_tile();
void fasta_index_tile(FastaTileArgs a, int n_sm, cudaStream_t s);

// ---- translate on wrapped FASTA read in place, proteins formatted straight into the output (k_translate_tile.cu)
struct TranslateTileArgs {
  const u8 *in;
  u32 n;
  const u8 *clean;                      // span flags of fasta_index_tile
  u32 n_rec, nf;
  int frames[8];
  const u32 *name_off, *name_len, *seq_start, *seq_len;  // record table
  u32 width_in, width_out;              // wrap width of the input lines (0 = one line) / of the proteins (0 = one line)
  unsigned long long magic_in, magic_out1;  // ceil(2^40 / width_in), ceil(2^40 / (width_out + 1))
  const u64 *out_off;                   // n_rec * nf + 1 element offsets
  u32 *tile_first;                      // scratch: translate_tile_tiles(total) entries
  u64 total;
  const u8 *code_fwd, *code_rev, *lut;  // IUPAC base codes (strand, complement strand), 4096-entry codon table
  const u8 *aa_fwd, *aa_rev;            // 64-entry tables for plain ACGT codons (clean applied), index = b0 | b1 << 2 | b2 << 4
  u64 start_fwd, start_rev;             // bit per 64-entry index: start codon (-M)
  int allow_unknown, init_m, clean_stop;
  u8 *out;
  DevStatus *st;
};
void translate_sizes(const u32 *name_len, const u32 *seq_len, u32 n_rec, u32 nf, const int *frames, u32 width_out, u32 *sizes,
                     DevStatus *st, cudaStream_t s);
u32 translate_tile_tiles(u64 total);
void translate_tile(const TranslateTileArgs &a, cudaStream_t s);

// ---- stats (k_stats.cu)
void stats_qual_gap(RecViews v, const u8 *gap, int fq_offset, int fastq, DevStatus *st, cudaStream_t s);

}  // namespace k
}  // namespace bsk
