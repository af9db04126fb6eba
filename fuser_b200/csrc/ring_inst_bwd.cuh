// ring_inst_bwd.cuh — instantiation table of the norm-backward ring kernels for one dtype.
#pragma once

#include "../../include/b200norm.h"
#include "kernels_bwd.cuh"
#include "ring_dispatch.cuh"

namespace b200norm {

template <typename T, bool LN>
RingKernel ring_bwd_kernel_sel(int tpr, int ch, bool cached) {
#define B200_BWD_CASE(TPR)                                                          \
  case TPR:                                                                         \
    if (cached) {                                                                   \
      if (ch == 1) return norm_bwd_ring_kernel<T, TPR, 1, LN, true>;                \
      if (ch == 2) return norm_bwd_ring_kernel<T, TPR, 2, LN, true>;                \
    } else {                                                                        \
      if (ch == 3) return norm_bwd_ring_kernel<T, TPR, 3, LN, false>;               \
      if (ch == 4 && !LN) return norm_bwd_ring_kernel<T, TPR, 4, false, false>;     \
    }                                                                               \
    return nullptr;
  switch (tpr) {
    B200_BWD_CASE(32)
    B200_BWD_CASE(64)
    B200_BWD_CASE(128)
    B200_BWD_CASE(256)
    B200_BWD_CASE(512)
    default: return nullptr;
  }
#undef B200_BWD_CASE
}

template <typename T>
RingKernel ring_bwd_kernel_for(int op, int tpr, int ch, bool cached) {
  switch (op) {
    case B200NORM_OP_LAYERNORM_BWD: return ring_bwd_kernel_sel<T, true>(tpr, ch, cached);
    case B200NORM_OP_RMSNORM_BWD: return ring_bwd_kernel_sel<T, false>(tpr, ch, cached);
    default: return nullptr;
  }
}

}  // namespace b200norm
