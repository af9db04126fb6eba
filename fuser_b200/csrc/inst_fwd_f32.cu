// Instantiations of the forward-type ring kernels for float (see ring_inst_fwd.cuh).
#include "ring_inst_fwd.cuh"
namespace b200norm {
RingKernel ring_fwd_kernel_f32(int op, int tpr, int ch, bool psmem) { return ring_fwd_kernel_for<float>(op, tpr, ch, psmem); }
}  // namespace b200norm
