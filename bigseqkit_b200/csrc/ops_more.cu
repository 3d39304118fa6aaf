// ops_more.cu -- RmDup, Translate, Locate, Grep (placeholders until implemented)
#include "engine.h"
#include "prims.h"

namespace bsk {
struct Engine::RmdupState {};
struct Engine::PatternSet {};
void Engine::free_op_state() {}
void Engine::reset_op_state() {}
int Engine::op_rmdup(BlockOut &, bool) { err = "rmdup not implemented"; return BSK_ERR_UNSUPPORTED; }
int Engine::op_translate(BlockOut &) { err = "translate not implemented"; return BSK_ERR_UNSUPPORTED; }
int Engine::op_locate(BlockOut &, int64_t) { err = "locate not implemented"; return BSK_ERR_UNSUPPORTED; }
int Engine::op_grep(BlockOut &) { err = "grep not implemented"; return BSK_ERR_UNSUPPORTED; }
int Engine::rmdup_keys(const int64_t **, size_t *) { err = "not implemented"; return BSK_ERR_UNSUPPORTED; }
int Engine::rmdup_prepare_device(const void *, size_t, void *, size_t, u64 *) { err = "not implemented"; return BSK_ERR_UNSUPPORTED; }
int Engine::rmdup_resolve_device(const void *, u64, bsk_out *) { err = "not implemented"; return BSK_ERR_UNSUPPORTED; }
}  // namespace bsk
