// Instantiations of the norm-backward ring kernels for float (see ring_inst_bwd.cuh).
#include "ring_inst_bwd.cuh"
namespace b200norm {
RingKernel ring_bwd_kernel_f32(int op, int tpr, int ch, bool cached, bool centered) {
  return ring_bwd_kernel_for<float>(op, tpr, ch, cached, centered);
}
RingKernel ring_bwd_drop_kernel_f32(int op, int tpr, int ch) { return ring_bwd_drop_kernel_for<float>(op, tpr, ch); }
RingKernel ring_bwd_rowaff_kernel_f32(int tpr, int ch, bool cached) { return ring_bwd_rowaff_kernel_for<float>(tpr, ch, cached); }
}  // namespace b200norm
