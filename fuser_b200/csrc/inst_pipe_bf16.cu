// Instantiations of the pipelined re-read backward for __nv_bfloat16 (see ring_inst_pipe.cuh).
#include "ring_inst_pipe.cuh"
namespace b200norm {
RingKernel ring_bwd_pipe_kernel_bf16(int op, int tpr, int ch, bool lean) { return ring_bwd_pipe_kernel_for<__nv_bfloat16>(op, tpr, ch, lean); }
RingKernel ring_fwd_pipe_kernel_bf16(int op, int tpr, int ch) { return ring_fwd_pipe_kernel_for<__nv_bfloat16>(op, tpr, ch); }
}  // namespace b200norm
