// Instantiations of the plain reductions on the ring for __half (see ring_inst_reduce.cuh).
#include "ring_inst_reduce.cuh"
namespace b200norm {
RingKernel ring_reduce_inner_kernel_f16(int tpr, int ch) { return ring_reduce_inner_kernel_for<__half>(tpr, ch); }
RingKernel ring_colsum_kernel_f16(int tpr, int ch) { return ring_colsum_kernel_for<__half>(tpr, ch); }
}  // namespace b200norm
