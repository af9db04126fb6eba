// Instantiations of the forward-type ring kernels for __nv_bfloat16 (see ring_inst_fwd.cuh).
#include "ring_inst_fwd.cuh"
namespace b200norm {
RingKernel ring_fwd_kernel_bf16(int op, int tpr, int ch, bool psmem) { return ring_fwd_kernel_for<__nv_bfloat16>(op, tpr, ch, psmem); }
}  // namespace b200norm
