// ring_inst_smpipe.cuh — instantiation table of the pipelined cluster softmax / log-softmax forward for one dtype.
#pragma once

#include "../../include/b200norm.h"
#include "kernels_softmax_pipe.cuh"
#include "ring_dispatch.cuh"

namespace b200norm {

// One 512-thread group per CTA (cluster plans only: slices of 16 KB and more); ch = chunks per thread kept packed in
// registers between the two halves of pass 1 (3 ... 8), or 0 = the looping form for wider slices.
template <typename T>
RingKernel ring_sm_pipe_kernel_for(int op, int tpr, int ch) {
  if (tpr != 512) return nullptr;
  const bool log = op == B200NORM_OP_LOGSOFTMAX_FWD;
  if (!log && op != B200NORM_OP_SOFTMAX_FWD) return nullptr;
#define B200_SMPIPE_CASE(CH) \
  case CH: return log ? softmax_fwd_pipe_kernel<T, 512, CH, true> : softmax_fwd_pipe_kernel<T, 512, CH, false>;
  switch (ch) {
    B200_SMPIPE_CASE(0)
    B200_SMPIPE_CASE(2)
    B200_SMPIPE_CASE(3)
    B200_SMPIPE_CASE(4)
    B200_SMPIPE_CASE(5)
    B200_SMPIPE_CASE(6)
    B200_SMPIPE_CASE(7)
    B200_SMPIPE_CASE(8)
    default: return nullptr;
  }
#undef B200_SMPIPE_CASE
}

}  // namespace b200norm
