// Instantiations of the pipelined one-round LayerNorm forward for __half (see ring_inst_lnpipe.cuh).
#include "ring_inst_lnpipe.cuh"
namespace b200norm {
RingKernel ring_ln_pipe_kernel_f16(int tpr, int ch, bool add) { return ring_ln_pipe_kernel_for<__half>(tpr, ch, add); }
}  // namespace b200norm
