// ops_more.cu -- RmDup, Translate, Locate, Grep (placeholders until implemented)
#include "engine.h"
#include "op_state.h"
#include "prims.h"

namespace bsk {
struct Engine::PatternSet {};
void Engine::free_op_state() {
  rmdup_state_free(rm_);
  rm_ = nullptr;
  delete pats_;
  pats_ = nullptr;
}
void Engine::reset_op_state() { rmdup_state_reset(rm_); }

int Engine::op_locate(BlockOut &, int64_t) { err = "locate not implemented"; return BSK_ERR_UNSUPPORTED; }
int Engine::op_grep(BlockOut &) { err = "grep not implemented"; return BSK_ERR_UNSUPPORTED; }
}  // namespace bsk
