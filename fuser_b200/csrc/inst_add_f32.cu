// Instantiations of the residual-add forward norm kernels for float (see ring_inst_add.cuh).
#include "ring_inst_add.cuh"
namespace b200norm {
RingKernel ring_add_kernel_f32(int op, int tpr, int ch, bool psmem) { return ring_add_kernel_for<float>(op, tpr, ch, psmem); }
}  // namespace b200norm
