// Instantiations of the residual-add forward norm kernels for __nv_bfloat16 (see ring_inst_add.cuh).
#include "ring_inst_add.cuh"
namespace b200norm {
RingKernel ring_add_kernel_bf16(int op, int tpr, int ch, bool psmem) { return ring_add_kernel_for<__nv_bfloat16>(op, tpr, ch, psmem); }
}  // namespace b200norm
