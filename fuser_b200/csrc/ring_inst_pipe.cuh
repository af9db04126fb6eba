// ring_inst_pipe.cuh — instantiation table of the pipelined re-read backward for one dtype.
#pragma once

#include "../../include/b200norm.h"
#include "kernels_bwd.cuh"
#include "ring_dispatch.cuh"

namespace b200norm {

template <typename T>
RingKernel ring_bwd_pipe_kernel_for(int op, int tpr, int ch) {
  const bool ln = op == B200NORM_OP_LAYERNORM_BWD;
  if (!ln && op != B200NORM_OP_RMSNORM_BWD) return nullptr;
#define B200_PIPE_CASE(TPR)                                                   \
  case TPR:                                                                   \
    if (ch == 3) return ln ? norm_bwd_pipe_kernel<T, TPR, 3, true> : norm_bwd_pipe_kernel<T, TPR, 3, false>; \
    if (ch == 4 && !ln) return norm_bwd_pipe_kernel<T, TPR, 4, false>;        \
    return nullptr;
  switch (tpr) {
    B200_PIPE_CASE(64)
    B200_PIPE_CASE(128)
    B200_PIPE_CASE(256)
    B200_PIPE_CASE(512)
    default: return nullptr;
  }
#undef B200_PIPE_CASE
}

}  // namespace b200norm
