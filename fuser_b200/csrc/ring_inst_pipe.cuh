// ring_inst_pipe.cuh — instantiation table of the pipelined re-read backward for one dtype.
#pragma once

#include "../../include/b200norm.h"
#include "kernels_bwd.cuh"
#include "kernels_bwd_lean.cuh"
#include "kernels_fwd.cuh"
#include "ring_dispatch.cuh"

namespace b200norm {

template <typename T>
RingKernel ring_bwd_pipe_kernel_for(int op, int tpr, int ch, bool lean) {
  const bool ln = op == B200NORM_OP_LAYERNORM_BWD;
  if (!ln && op != B200NORM_OP_RMSNORM_BWD) return nullptr;
  if (lean && !ln) {  // RMSNorm backward without the centring term: the lean kernel (kernels_bwd_lean.cuh)
    if (tpr == 512 && ch == 4) return rms_bwd_pipe_lean_kernel<T, 512, 4>;
    if (tpr == 512 && ch == 3) return rms_bwd_pipe_lean_kernel<T, 512, 3>;
  }
#define B200_PIPE_CASE(TPR)                                                   \
  case TPR:                                                                   \
    if (ch == 3) return ln ? norm_bwd_pipe_kernel<T, TPR, 3, true> : norm_bwd_pipe_kernel<T, TPR, 3, false>; \
    if (ch == 4 && !ln) return norm_bwd_pipe_kernel<T, TPR, 4, false>;        \
    return nullptr;
  // (the re-read form is only chosen when 512 threads x 2 chunks do not hold the slice: 512 threads per row)
  switch (tpr) {
#ifdef B200NORM_ALL_KERNELS
    B200_PIPE_CASE(64)
    B200_PIPE_CASE(128)
    B200_PIPE_CASE(256)
#endif
    B200_PIPE_CASE(512)
    default: return nullptr;
  }
#undef B200_PIPE_CASE
}

// RMSNorm forward, pipelined (cluster rows): ch in {1..4} (tpr 64 / 128), {1..6} (tpr 256 / 512)
template <typename T>
RingKernel ring_fwd_pipe_kernel_for(int op, int tpr, int ch) {
  if (op != B200NORM_OP_RMSNORM_FWD) return nullptr;
#ifdef B200NORM_ALL_KERNELS
#define B200_FPIPE_LOW(TPR)                                                   \
    case 1: return rmsnorm_fwd_pipe_kernel<T, TPR, 1>;                        \
    case 2: return rmsnorm_fwd_pipe_kernel<T, TPR, 2>;
#else
#define B200_FPIPE_LOW(TPR)
#endif
#define B200_FPIPE_CH(TPR, MAXCH)                                              \
  switch (ch) {                                                               \
    B200_FPIPE_LOW(TPR)                                                       \
    case 3: return rmsnorm_fwd_pipe_kernel<T, TPR, 3>;                        \
    case 4: return rmsnorm_fwd_pipe_kernel<T, TPR, 4>;                        \
    case 5: if (MAXCH >= 5) return rmsnorm_fwd_pipe_kernel<T, TPR, (MAXCH >= 5 ? 5 : 1)>; return nullptr; \
    case 6: if (MAXCH >= 6) return rmsnorm_fwd_pipe_kernel<T, TPR, (MAXCH >= 6 ? 6 : 1)>; return nullptr; \
    default: return nullptr;                                                  \
  }
  switch (tpr) {
    case 64: B200_FPIPE_CH(64, 4)
    case 128: B200_FPIPE_CH(128, 4)
    case 256: B200_FPIPE_CH(256, 6)
    case 512: B200_FPIPE_CH(512, 6)
    default: return nullptr;
  }
#undef B200_FPIPE_CH
#undef B200_FPIPE_LOW
}

}  // namespace b200norm
