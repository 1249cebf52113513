// srw_stubs — entry points declared in include/srw.h whose kernels are not built yet.  They fail loudly
// (SRW_ERR_UNSUPPORTED + message); nothing here computes anything and there is no fallback behind them.
#include "../../include/srw.h"
#include "srw_common.cuh"

#define SRW_STUB(name, ...)                                         \
  extern "C" int name(__VA_ARGS__) {                                \
    ::srw::set_last_error(#name ": not implemented in this build"); \
    return SRW_ERR_UNSUPPORTED;                                     \
  }

SRW_STUB(srw_rewarder_fwd, const srw_rewarder_fwd_args*, void*)
SRW_STUB(srw_generator_fwd, const srw_generator_fwd_args*, void*)
SRW_STUB(srw_rewarder_train, const srw_rewarder_train_args*, void*)
SRW_STUB(srw_flexmatch_epilogue, const srw_flexmatch_epilogue_args*, void*)
SRW_STUB(srw_adamw_step, const srw_adamw_args*, void*)
extern "C" int64_t srw_rewarder_workspace_floats(int, int) { return -1; }
extern "C" int64_t srw_rewarder_train_workspace_floats(int, int, int) { return -1; }
extern "C" int64_t srw_adamw_table_bytes(int) { return -1; }
