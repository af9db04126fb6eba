// ring_inst_reduce.cuh — instantiation table of the plain reductions on the ring for one dtype.
#pragma once

#include "../../include/b200norm.h"
#include "kernels_reduce.cuh"
#include "ring_dispatch.cuh"

namespace b200norm {

#define B200_RED_CH(K, TPR)                    \
  switch (ch) {                                \
    case 1: return K<T, TPR, 1>;               \
    case 2: return K<T, TPR, 2>;               \
    case 3: return K<T, TPR, 3>;               \
    case 4: return K<T, TPR, 4>;               \
    default: return nullptr;                   \
  }
#ifdef B200NORM_ALL_KERNELS
#define B200_RED_CH_HI B200_RED_CH
#else  // from 64 threads per row on the planner only emits 3 or 4 chunks per thread (see ring_inst_fwd.cuh)
#define B200_RED_CH_HI(K, TPR)                 \
  switch (ch) {                                \
    case 3: return K<T, TPR, 3>;               \
    case 4: return K<T, TPR, 4>;               \
    default: return nullptr;                   \
  }
#endif
#define B200_RED_TPR(K)                        \
  switch (tpr) {                               \
    case 32: B200_RED_CH(K, 32)                \
    case 64: B200_RED_CH_HI(K, 64)             \
    case 128: B200_RED_CH_HI(K, 128)           \
    case 256: B200_RED_CH_HI(K, 256)           \
    case 512: B200_RED_CH_HI(K, 512)           \
    default: return nullptr;                   \
  }

template <typename T>
RingKernel ring_reduce_inner_kernel_for(int tpr, int ch) {
  if (tpr == 512 && ch == 0) return reduce_inner_ring_kernel<T, 512, 0>;
  if (tpr == 256 && ch == 0) return reduce_inner_ring_kernel<T, 256, 0>;
  if (tpr == 128 && ch == 0) return reduce_inner_ring_kernel<T, 128, 0>;
  B200_RED_TPR(reduce_inner_ring_kernel)
}
template <typename T>
RingKernel ring_colsum_kernel_for(int tpr, int ch) {
  B200_RED_TPR(colsum_ring_kernel)
}
#undef B200_RED_TPR
#undef B200_RED_CH_HI
#undef B200_RED_CH

}  // namespace b200norm
