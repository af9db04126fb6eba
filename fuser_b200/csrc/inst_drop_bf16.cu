// Instantiations of the forward norm kernels with the fused dropout + residual add for __nv_bfloat16 (see ring_inst_add.cuh).
#include "ring_inst_add.cuh"
namespace b200norm {
RingKernel ring_drop_kernel_bf16(int op, int tpr, int ch) { return ring_drop_kernel_for<__nv_bfloat16>(op, tpr, ch); }
}  // namespace b200norm
