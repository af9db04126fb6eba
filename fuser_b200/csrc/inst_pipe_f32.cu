// Instantiations of the pipelined re-read backward for float (see ring_inst_pipe.cuh).
#include "ring_inst_pipe.cuh"
namespace b200norm {
RingKernel ring_bwd_pipe_kernel_f32(int op, int tpr, int ch, bool lean) { return ring_bwd_pipe_kernel_for<float>(op, tpr, ch, lean); }
RingKernel ring_fwd_pipe_kernel_f32(int op, int tpr, int ch) { return ring_fwd_pipe_kernel_for<float>(op, tpr, ch); }
}  // namespace b200norm
