// Instantiations of the pipelined one-round LayerNorm forward for float (see ring_inst_lnpipe.cuh).
#include "ring_inst_lnpipe.cuh"
namespace b200norm {
RingKernel ring_ln_pipe_kernel_f32(int tpr, int ch, bool add) { return ring_ln_pipe_kernel_for<float>(tpr, ch, add); }
}  // namespace b200norm
