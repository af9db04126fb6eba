// ring_inst_lnpipe.cuh — instantiation table of the pipelined one-round LayerNorm forward for one dtype.
#pragma once

#include "../../include/b200norm.h"
#include "kernels_ln_pipe.cuh"
#include "ring_dispatch.cuh"

namespace b200norm {

template <typename T>
RingKernel ring_ln_pipe_kernel_for(int tpr, int ch, bool add) {
// (3 or 4 chunks per thread: what the planner emits from 64 threads per row on, see ring_inst_fwd.cuh)
#ifdef B200NORM_ALL_KERNELS
#define B200_LNPIPE_LOW(TPR, ADD)                                     \
    case 1: return layernorm_fwd_pipe_kernel<T, TPR, 1, ADD>;         \
    case 2: return layernorm_fwd_pipe_kernel<T, TPR, 2, ADD>;
#else
#define B200_LNPIPE_LOW(TPR, ADD)
#endif
#define B200_LNPIPE_CH(TPR, ADD)                                      \
  switch (ch) {                                                       \
    B200_LNPIPE_LOW(TPR, ADD)                                         \
    case 3: return layernorm_fwd_pipe_kernel<T, TPR, 3, ADD>;         \
    case 4: return layernorm_fwd_pipe_kernel<T, TPR, 4, ADD>;         \
    default: return nullptr;                                          \
  }
  if (add) {
    switch (tpr) {
      case 64: B200_LNPIPE_CH(64, true)
      case 128: B200_LNPIPE_CH(128, true)
      case 256: B200_LNPIPE_CH(256, true)
      case 512: B200_LNPIPE_CH(512, true)
      default: return nullptr;
    }
  }
  // two 256-thread groups per CTA for rows of 5 / 6 chunks per thread (N = 10240 / 12288 at 16 bits)
  if (tpr == 256 && ch == 5) return layernorm_fwd_pipe_kernel<T, 256, 5, false>;
  if (tpr == 256 && ch == 6) return layernorm_fwd_pipe_kernel<T, 256, 6, false>;
  // ... and one 512-thread group for rows of 5 / 6 chunks per thread (N = 20480 / 24576 at 16 bits) without a cluster
  if (tpr == 512 && ch == 5) return layernorm_fwd_pipe_kernel<T, 512, 5, false>;
  if (tpr == 512 && ch == 6) return layernorm_fwd_pipe_kernel<T, 512, 6, false>;
  switch (tpr) {
    case 64: B200_LNPIPE_CH(64, false)
    case 128: B200_LNPIPE_CH(128, false)
    case 256: B200_LNPIPE_CH(256, false)
    case 512: B200_LNPIPE_CH(512, false)
    default: return nullptr;
  }
#undef B200_LNPIPE_CH
#undef B200_LNPIPE_LOW
}

}  // namespace b200norm
