// ring_inst_add.cuh — instantiation table of the forward norm kernels with the fused
// residual add (norm(x + residual), optional sum output) for one dtype.
#pragma once

#include "../../include/b200norm.h"
#include "kernels_fwd.cuh"
#include "ring_dispatch.cuh"

namespace b200norm {

#define B200_ADD_CH_SWITCH(K, TPR, ...)                 \
  switch (ch) {                                         \
    case 1: return K<T, TPR, 1, ##__VA_ARGS__>;         \
    case 2: return K<T, TPR, 2, ##__VA_ARGS__>;         \
    case 3: return K<T, TPR, 3, ##__VA_ARGS__>;         \
    case 4: return K<T, TPR, 4, ##__VA_ARGS__>;         \
    default: return nullptr;                            \
  }
#ifdef B200NORM_ALL_KERNELS
#define B200_ADD_CH_SWITCH_HI B200_ADD_CH_SWITCH
#else  // from 64 threads per row on the planner only emits 3 or 4 chunks per thread (see ring_inst_fwd.cuh)
#define B200_ADD_CH_SWITCH_HI(K, TPR, ...)              \
  switch (ch) {                                         \
    case 3: return K<T, TPR, 3, ##__VA_ARGS__>;         \
    case 4: return K<T, TPR, 4, ##__VA_ARGS__>;         \
    default: return nullptr;                            \
  }
#endif
#define B200_ADD_TPR_SWITCH(K, ...)                                       \
  switch (tpr) {                                                          \
    case 32: B200_ADD_CH_SWITCH(K, 32, ##__VA_ARGS__)                     \
    case 64: B200_ADD_CH_SWITCH_HI(K, 64, ##__VA_ARGS__)                  \
    case 128: B200_ADD_CH_SWITCH_HI(K, 128, ##__VA_ARGS__)                \
    case 256: B200_ADD_CH_SWITCH_HI(K, 256, ##__VA_ARGS__)                \
    case 512:                                                             \
      if (ch == 0) return K<T, 512, 0, ##__VA_ARGS__>;                    \
      B200_ADD_CH_SWITCH_HI(K, 512, ##__VA_ARGS__)                        \
    default: return nullptr;                                              \
  }

// (the dropout kernels take 1..4 chunks per thread at every group size: the planner picks <= 2 for short rows)
#define B200_ADD_TPR_SWITCH_NOLOOP(K, ...)                                \
  switch (tpr) {                                                          \
    case 32: B200_ADD_CH_SWITCH(K, 32, ##__VA_ARGS__)                     \
    case 64: B200_ADD_CH_SWITCH(K, 64, ##__VA_ARGS__)                     \
    case 128: B200_ADD_CH_SWITCH(K, 128, ##__VA_ARGS__)                   \
    case 256: B200_ADD_CH_SWITCH(K, 256, ##__VA_ARGS__)                   \
    case 512: B200_ADD_CH_SWITCH(K, 512, ##__VA_ARGS__)                   \
    default: return nullptr;                                              \
  }

template <typename T>
RingKernel ring_add_kernel_for(int op, int tpr, int ch, bool psmem) {
  switch (op) {
    case B200NORM_OP_RMSNORM_FWD: B200_ADD_TPR_SWITCH(rmsnorm_fwd_ring_kernel, true)
    case B200NORM_OP_LAYERNORM_FWD:
      if (psmem && ch > 0) {
        B200_ADD_TPR_SWITCH(layernorm_fwd_ring_kernel, true, true)
      }
      B200_ADD_TPR_SWITCH(layernorm_fwd_ring_kernel, false, true)
    default: return nullptr;
  }
}
// ... with the fused dropout of the x stream (register-resident form, parameters in registers)
template <typename T>
RingKernel ring_drop_kernel_for(int op, int tpr, int ch) {
  if (ch <= 0) return nullptr;
  switch (op) {
    case B200NORM_OP_RMSNORM_FWD: B200_ADD_TPR_SWITCH_NOLOOP(rmsnorm_fwd_ring_kernel, true, true)
    case B200NORM_OP_LAYERNORM_FWD: B200_ADD_TPR_SWITCH_NOLOOP(layernorm_fwd_ring_kernel, false, true, true)
    default: return nullptr;
  }
}
#undef B200_ADD_TPR_SWITCH_NOLOOP
#undef B200_ADD_TPR_SWITCH
#undef B200_ADD_CH_SWITCH_HI
#undef B200_ADD_CH_SWITCH

}  // namespace b200norm
