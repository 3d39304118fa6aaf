// op_state.h -- operator state that outlives one block (rmdup history + table, pattern tables).
#pragma once
#include <string>
#include <vector>

#include "engine.h"

namespace bsk {

struct Engine::RmdupState {
  // key -> earliest global record ordinal; open addressing, cap slots + 1 extra slot for key 0
  u64 *tkeys = nullptr, *tfirst = nullptr;
  u64 cap = 0;
  // {xxh64 seed 0, xxh64 seed B} of every record of the partition seen so far (global input order)
  u64 *hist_keys = nullptr, *hist_fp = nullptr;
  u64 n_hist = 0, hist_cap = 0;
  // subject slices of the block in flight
  const u8 *sv_base = nullptr;
  const u32 *sv_off = nullptr, *sv_len = nullptr;
  bool block_ready = false;
};

void rmdup_state_free(Engine::RmdupState *rm);
void rmdup_state_reset(Engine::RmdupState *rm);

}  // namespace bsk
