// Instantiations of the pipelined one-round LayerNorm forward for __nv_bfloat16 (see ring_inst_lnpipe.cuh).
#include "ring_inst_lnpipe.cuh"
namespace b200norm {
RingKernel ring_ln_pipe_kernel_bf16(int tpr, int ch, bool add) { return ring_ln_pipe_kernel_for<__nv_bfloat16>(tpr, ch, add); }
}  // namespace b200norm
