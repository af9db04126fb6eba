// ring_inst_smpipe.cuh — instantiation table of the pipelined cluster softmax / log-softmax / cross-entropy forward for one dtype.
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
  // REDUCE_INNER here means its cross-entropy flavour (B200NORM_F_REDUCE_NLL): the planner sets pipe for no other
  const int mode = op == B200NORM_OP_SOFTMAX_FWD ? 0 : op == B200NORM_OP_LOGSOFTMAX_FWD ? 1 : op == B200NORM_OP_REDUCE_INNER ? 2 : -1;
  if (mode < 0) return nullptr;
#define B200_SMPIPE_CASE(CH)                                           \
  case CH:                                                             \
    return mode == 0   ? softmax_fwd_pipe_kernel<T, 512, CH, 0>        \
           : mode == 1 ? softmax_fwd_pipe_kernel<T, 512, CH, 1>        \
                       : softmax_fwd_pipe_kernel<T, 512, CH, 2>;
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
