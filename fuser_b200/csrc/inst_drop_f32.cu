// Instantiations of the forward norm kernels with the fused dropout + residual add for float (see ring_inst_add.cuh).
#include "ring_inst_add.cuh"
namespace b200norm {
RingKernel ring_drop_kernel_f32(int op, int tpr, int ch) { return ring_drop_kernel_for<float>(op, tpr, ch); }
}  // namespace b200norm
