// Instantiations of the pipelined cluster softmax forward for __half (see ring_inst_smpipe.cuh).
#include "ring_inst_smpipe.cuh"
namespace b200norm {
RingKernel ring_sm_pipe_kernel_f16(int op, int tpr, int ch) { return ring_sm_pipe_kernel_for<__half>(op, tpr, ch); }
}  // namespace b200norm
