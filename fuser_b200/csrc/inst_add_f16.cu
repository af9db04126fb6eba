// Instantiations of the residual-add forward norm kernels for __half (see ring_inst_add.cuh).
#include "ring_inst_add.cuh"
namespace b200norm {
RingKernel ring_add_kernel_f16(int op, int tpr, int ch, bool psmem) { return ring_add_kernel_for<__half>(op, tpr, ch, psmem); }
}  // namespace b200norm
