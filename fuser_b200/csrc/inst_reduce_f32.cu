// Instantiations of the plain reductions on the ring for float (see ring_inst_reduce.cuh).
#include "ring_inst_reduce.cuh"
namespace b200norm {
RingKernel ring_reduce_inner_kernel_f32(int tpr, int ch) { return ring_reduce_inner_kernel_for<float>(tpr, ch); }
RingKernel ring_colsum_kernel_f32(int tpr, int ch) { return ring_colsum_kernel_for<float>(tpr, ch); }
}  // namespace b200norm
