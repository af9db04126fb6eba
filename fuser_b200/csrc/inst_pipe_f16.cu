// Instantiations of the pipelined re-read backward for __half (see ring_inst_pipe.cuh).
#include "ring_inst_pipe.cuh"
namespace b200norm {
RingKernel ring_bwd_pipe_kernel_f16(int op, int tpr, int ch, bool lean) { return ring_bwd_pipe_kernel_for<__half>(op, tpr, ch, lean); }
RingKernel ring_fwd_pipe_kernel_f16(int op, int tpr, int ch) { return ring_fwd_pipe_kernel_for<__half>(op, tpr, ch); }
}  // namespace b200norm
