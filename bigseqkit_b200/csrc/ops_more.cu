// ops_more.cu -- RmDup, Translate, Locate, Grep (placeholders until implemented)
#include "engine.h"
#include "op_state.h"
#include "prims.h"

namespace bsk {
void Engine::free_op_state() {
  rmdup_state_free(rm_);
  rm_ = nullptr;
  delete pats_;
  pats_ = nullptr;
}
void Engine::reset_op_state() { rmdup_state_reset(rm_); }

}  // namespace bsk
