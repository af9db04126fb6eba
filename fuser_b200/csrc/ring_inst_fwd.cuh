// ring_inst_fwd.cuh — instantiation table of the forward-type ring kernels for one dtype.
#pragma once

#include "../../include/b200norm.h"
#include "kernels_fwd.cuh"
#include "ring_dispatch.cuh"

namespace b200norm {

// Only the (threads per row, chunks per thread) pairs the planner can emit are instantiated: it picks the SMALLEST
// group size whose threads hold the slice at <= 4 chunks each, so from 64 threads on a thread owns 3 or 4 chunks
// (1 or 2 would have fit the next smaller group). -DB200NORM_ALL_KERNELS restores the full cross product for
// tuning with B200NORM_TPR.
#ifdef B200NORM_ALL_KERNELS
#define B200_LOW_CH(K, TPR, ...)                        \
    case 1: return K<T, TPR, 1, ##__VA_ARGS__>;         \
    case 2: return K<T, TPR, 2, ##__VA_ARGS__>;
#else
#define B200_LOW_CH(K, TPR, ...)
#endif
#define B200_CH_SWITCH(K, TPR, ...)                     \
  switch (ch) {                                         \
    case 1: return K<T, TPR, 1, ##__VA_ARGS__>;         \
    case 2: return K<T, TPR, 2, ##__VA_ARGS__>;         \
    case 3: return K<T, TPR, 3, ##__VA_ARGS__>;         \
    case 4: return K<T, TPR, 4, ##__VA_ARGS__>;         \
    default: return nullptr;                            \
  }
#define B200_CH_SWITCH_HI(K, TPR, ...)                  \
  switch (ch) {                                         \
    B200_LOW_CH(K, TPR, ##__VA_ARGS__)                  \
    case 3: return K<T, TPR, 3, ##__VA_ARGS__>;         \
    case 4: return K<T, TPR, 4, ##__VA_ARGS__>;         \
    default: return nullptr;                            \
  }
// CH 5 / 6 (TPR 256 / 512 only): kernels whose per-chunk register state is just x keep two
// consumer groups per CTA at N = 10240 / 12288 instead of one 512-thread group.
#define B200_CH_SWITCH_WIDE(K, TPR, ...)                \
  switch (ch) {                                         \
    B200_LOW_CH(K, TPR, ##__VA_ARGS__)                  \
    case 3: return K<T, TPR, 3, ##__VA_ARGS__>;         \
    case 4: return K<T, TPR, 4, ##__VA_ARGS__>;         \
    case 5: return K<T, TPR, 5, ##__VA_ARGS__>;         \
    case 6: return K<T, TPR, 6, ##__VA_ARGS__>;         \
    default: return nullptr;                            \
  }
#define B200_TPR_SWITCH_WIDE(K, ...)                                      \
  switch (tpr) {                                                          \
    case 32: B200_CH_SWITCH(K, 32, ##__VA_ARGS__)                         \
    case 64: B200_CH_SWITCH_HI(K, 64, ##__VA_ARGS__)                      \
    case 128: B200_CH_SWITCH_HI(K, 128, ##__VA_ARGS__)                    \
    case 256: B200_CH_SWITCH_WIDE(K, 256, ##__VA_ARGS__)                  \
    case 512:                                                             \
      if (ch == 0) return K<T, 512, 0, ##__VA_ARGS__>;                    \
      B200_CH_SWITCH_WIDE(K, 512, ##__VA_ARGS__)                          \
    default: return nullptr;                                              \
  }
#define B200_TPR_SWITCH(K, ...)                                           \
  switch (tpr) {                                                          \
    case 32: B200_CH_SWITCH(K, 32, ##__VA_ARGS__)                         \
    case 64: B200_CH_SWITCH_HI(K, 64, ##__VA_ARGS__)                      \
    case 128: B200_CH_SWITCH_HI(K, 128, ##__VA_ARGS__)                    \
    case 256: B200_CH_SWITCH_HI(K, 256, ##__VA_ARGS__)                    \
    case 512:                                                             \
      if (ch == 0) return K<T, 512, 0, ##__VA_ARGS__>;                    \
      B200_CH_SWITCH_HI(K, 512, ##__VA_ARGS__)                            \
    default: return nullptr;                                              \
  }

template <typename T>
RingKernel ring_fwd_kernel_for(int op, int tpr, int ch, bool psmem) {
  switch (op) {
    case B200NORM_OP_RMSNORM_FWD: B200_TPR_SWITCH_WIDE(rmsnorm_fwd_ring_kernel)
    case B200NORM_OP_LAYERNORM_FWD:
      if (psmem && ch > 0) {
        B200_TPR_SWITCH(layernorm_fwd_ring_kernel, true)
      }
      B200_TPR_SWITCH(layernorm_fwd_ring_kernel)
    case B200NORM_OP_SOFTMAX_FWD: B200_TPR_SWITCH_WIDE(softmax_fwd_ring_kernel, false)
    case B200NORM_OP_LOGSOFTMAX_FWD: B200_TPR_SWITCH(softmax_fwd_ring_kernel, true)
    case B200NORM_OP_SOFTMAX_BWD: B200_TPR_SWITCH(softmax_bwd_ring_kernel, false)
    case B200NORM_OP_LOGSOFTMAX_BWD: B200_TPR_SWITCH(softmax_bwd_ring_kernel, true)
    default: return nullptr;
  }
}
#undef B200_TPR_SWITCH
#undef B200_TPR_SWITCH_WIDE
#undef B200_CH_SWITCH_WIDE
#undef B200_CH_SWITCH_HI
#undef B200_CH_SWITCH
#undef B200_LOW_CH

}  // namespace b200norm
