// op_state.h -- operator state that outlives one block (rmdup history + table, pattern tables).
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "engine.h"

namespace bsk {

struct __align__(16) TableSlot {
  u64 key, first;
};

struct Engine::RmdupState {
  // key -> earliest global record ordinal; open addressing, cap slots + 1 extra slot for the all-ones key
  TableSlot *table = nullptr;
  u64 cap = 0, alloc_cap = 0;
  bool dirty = false;  // holds keys of a partition that has been reset
  // {xxh64 seed 0, xxh64 seed B} of every record of the partition seen so far (global input order)
  u64 *hist_keys = nullptr, *hist_fp = nullptr;
  u64 n_hist = 0, hist_cap = 0;
  // subject slices of the block in flight
  const u8 *sv_base = nullptr;
  const u32 *sv_off = nullptr, *sv_len = nullptr;
  u64 sv_limit = 0;
  bool block_ready = false;
  // rmdup -d / -D, accumulated on the host between Before and After (bigseqkit-lib/rmdup.go:102-103,224-238)
  std::string dup_seqs;                        // removed records, Record.Format(LineWidth), input order
  std::string id_text;                         // "ID\n" of every record seen (-D only)
  std::vector<u64> id_off;                     // record ordinal -> offset into id_text
  std::vector<std::pair<u64, u64>> dup_pairs;  // {ordinal of the group's first member, ordinal of the removed record}
  std::string dup_num;                         // rendered rows
};

// Needles of locate / grep -s: every pattern on the '+' strand and, unless only the positive
// strand is searched, reverse(pair(pattern)) which is matched on the forward strand instead of
// matching the pattern on a reverse-complemented copy of the sequence (bigseqkit-lib/locate.go:669-766).
struct Engine::PatternSet {
  int alphabet = -1;
  bool only_pos = false;
  u32 n_needles = 0, n_groups = 0, max_len = 0;
  // device arrays (one allocation each)
  DevBuf bytes;      // needle bytes arena
  DevBuf meta;       // u32 x 4 per needle: byte offset, length, (pattern << 1 | strand), hash  (sorted by group, hash)
  DevBuf groups;     // u32 x 4 per length group: length, first needle, needle count, table offset (u32 units)
  DevBuf tables;     // per group: u32 table_size followed by table_size slots (needle id + 1, 0 = empty)
  // locate rows: pattern names / pattern text
  DevBuf pat_bytes;  // names then patterns
  DevBuf pat_meta;   // u32 x 4 per pattern: name offset, name len, pattern offset, pattern len
  // grep by id / name: sorted 64-bit hashes + pattern slices
  DevBuf name_hash;  // u64 per pattern (sorted)
  DevBuf name_meta;  // u32 x 2 per pattern: offset, len into pat_bytes (same order as name_hash)
  u32 n_names = 0;
  // host copy of the needles (group order) and, for an equal-length ACGT panel, the 2-bit tables of k_locate_tile.cu
  std::vector<std::string> needle_s;
  std::vector<u32> needle_ps;
  bool kmer_built = false, kmer_ok = false;
  u32 kL = 0, kfbits = 0, kmul = 0, kmul2 = 0, kcmask = 0, ktmask = 0, ktshift = 0, kn = 0, kvmask = 0, kvbase = 0;
  DevBuf klut, kfilter, kfptab, ktable, kcode, kps;
};

void rmdup_state_free(Engine::RmdupState *rm);
void rmdup_state_reset(Engine::RmdupState *rm);

}  // namespace bsk
