// Instantiations of the forward-type ring kernels for __half (see ring_inst_fwd.cuh).
#include "ring_inst_fwd.cuh"
namespace b200norm {
RingKernel ring_fwd_kernel_f16(int op, int tpr, int ch, bool psmem) { return ring_fwd_kernel_for<__half>(op, tpr, ch, psmem); }
}  // namespace b200norm
