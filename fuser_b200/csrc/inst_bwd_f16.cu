// Instantiations of the norm-backward ring kernels for __half (see ring_inst_bwd.cuh).
#include "ring_inst_bwd.cuh"
namespace b200norm {
RingKernel ring_bwd_kernel_f16(int op, int tpr, int ch, bool cached, bool centered) {
  return ring_bwd_kernel_for<__half>(op, tpr, ch, cached, centered);
}
RingKernel ring_bwd_drop_kernel_f16(int op, int tpr, int ch) { return ring_bwd_drop_kernel_for<__half>(op, tpr, ch); }
RingKernel ring_bwd_rowaff_kernel_f16(int tpr, int ch, bool cached) { return ring_bwd_rowaff_kernel_for<__half>(tpr, ch, cached); }
}  // namespace b200norm
