// ring_inst_bwd.cuh — instantiation table of the norm-backward ring kernels for one dtype.
#pragma once

#include "../../include/b200norm.h"
#include "kernels_bwd.cuh"
#include "ring_dispatch.cuh"

namespace b200norm {

// Only what the planner can emit is instantiated (-DB200NORM_ALL_KERNELS: the full cross product, for tuning with
// B200NORM_TPR / B200NORM_BWD_REREAD): the cached form takes the smallest group that holds the slice at <= 2 chunks
// per thread — so from 64 threads on it is always 2 —, and the re-read form is only chosen when 512 threads x 2 chunks
// do not hold the slice, i.e. at 512 threads per row.
template <typename T, bool LN, bool CEN>
RingKernel ring_bwd_kernel_sel(int tpr, int ch, bool cached) {
#ifdef B200NORM_ALL_KERNELS
#define B200_BWD_CASE(TPR, LOW, REREAD)                                             \
  case TPR:                                                                         \
    if (cached) {                                                                   \
      if (ch == 1) return norm_bwd_ring_kernel<T, TPR, 1, LN, true, CEN>;                \
      if (ch == 2) return norm_bwd_ring_kernel<T, TPR, 2, LN, true, CEN>;                \
    } else {                                                                        \
      if (ch == 3) return norm_bwd_ring_kernel<T, TPR, 3, LN, false, CEN>;               \
      if (ch == 4 && !LN) return norm_bwd_ring_kernel<T, TPR, 4, false, false, CEN>;     \
    }                                                                               \
    return nullptr;
#else
#define B200_BWD_CASE(TPR, LOW, REREAD)                                             \
  case TPR:                                                                         \
    if (cached) {                                                                   \
      if (LOW && ch == 1) return norm_bwd_ring_kernel<T, TPR, LOW ? 1 : 2, LN, true, CEN>; \
      if (ch == 2) return norm_bwd_ring_kernel<T, TPR, 2, LN, true, CEN>;                \
    } else if (REREAD) {                                                            \
      if (ch == 3) return norm_bwd_ring_kernel<T, TPR, 3, LN, !REREAD, CEN>;             \
      if (ch == 4 && !LN) return norm_bwd_ring_kernel<T, TPR, REREAD ? 4 : 3, false, !REREAD, CEN>; \
    }                                                                               \
    return nullptr;
#endif
  switch (tpr) {
    B200_BWD_CASE(32, true, false)
    B200_BWD_CASE(64, false, false)
    B200_BWD_CASE(128, false, false)
    B200_BWD_CASE(256, false, false)
    B200_BWD_CASE(512, false, true)
    default: return nullptr;
  }
#undef B200_BWD_CASE
}

// centered: the dx formula keeps -sum(g) (always for LayerNorm; B200NORM_F_RMS_BWD_CENTERED for RMSNorm)
template <typename T>
RingKernel ring_bwd_kernel_for(int op, int tpr, int ch, bool cached, bool centered) {
  switch (op) {
    case B200NORM_OP_LAYERNORM_BWD: return ring_bwd_kernel_sel<T, true, true>(tpr, ch, cached);
    case B200NORM_OP_RMSNORM_BWD:
      // (the textbook specialisation exists for the re-read form only: the cached form measured slower with it)
      return (centered || cached) ? ring_bwd_kernel_sel<T, false, true>(tpr, ch, cached)
                                  : ring_bwd_kernel_sel<T, false, false>(tpr, ch, false);
    default: return nullptr;
  }
}

// fused-dropout backward (cached form; the (tpr, ch) pairs of the cached table above)
template <typename T, bool LN>
RingKernel ring_bwd_drop_kernel_sel(int tpr, int ch) {
  switch (tpr) {
    case 32:
      if (ch == 1) return norm_bwd_ring_kernel<T, 32, 1, LN, true, true, true>;
      if (ch == 2) return norm_bwd_ring_kernel<T, 32, 2, LN, true, true, true>;
      return nullptr;
    case 64: return ch == 2 ? norm_bwd_ring_kernel<T, 64, 2, LN, true, true, true> : nullptr;
    case 128: return ch == 2 ? norm_bwd_ring_kernel<T, 128, 2, LN, true, true, true> : nullptr;
    case 256: return ch == 2 ? norm_bwd_ring_kernel<T, 256, 2, LN, true, true, true> : nullptr;
    case 512: return ch == 2 ? norm_bwd_ring_kernel<T, 512, 2, LN, true, true, true> : nullptr;
    default: return nullptr;
  }
}
// instance norm backward (LayerNorm with the row's channel weight, per-row sums out): the cached pairs and the
// re-read form at 512 threads
template <typename T>
RingKernel ring_bwd_rowaff_kernel_for(int tpr, int ch, bool cached) {
  if (cached) {
    switch (tpr) {
      case 32:
        if (ch == 1) return norm_bwd_ring_kernel<T, 32, 1, true, true, true, false, true>;
        if (ch == 2) return norm_bwd_ring_kernel<T, 32, 2, true, true, true, false, true>;
        return nullptr;
      case 64: return ch == 2 ? norm_bwd_ring_kernel<T, 64, 2, true, true, true, false, true> : nullptr;
      case 128: return ch == 2 ? norm_bwd_ring_kernel<T, 128, 2, true, true, true, false, true> : nullptr;
      case 256: return ch == 2 ? norm_bwd_ring_kernel<T, 256, 2, true, true, true, false, true> : nullptr;
      case 512: return ch == 2 ? norm_bwd_ring_kernel<T, 512, 2, true, true, true, false, true> : nullptr;
      default: return nullptr;
    }
  }
  return (tpr == 512 && ch == 3) ? norm_bwd_ring_kernel<T, 512, 3, true, false, true, false, true> : nullptr;
}
template <typename T>
RingKernel ring_bwd_drop_kernel_for(int op, int tpr, int ch) {
  return op == B200NORM_OP_LAYERNORM_BWD ? ring_bwd_drop_kernel_sel<T, true>(tpr, ch)
                                         : ring_bwd_drop_kernel_sel<T, false>(tpr, ch);
}

}  // namespace b200norm
