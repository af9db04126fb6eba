// ring_dispatch.cuh — lookup of ring-kernel instantiations. The instantiations live in
// one translation unit per (dtype, direction) so the library builds in parallel
// (inst_fwd_*.cu / inst_bwd_*.cu); this header is the only thing b200norm.cu sees.
#pragma once

#include "ring.cuh"

namespace b200norm {

using RingKernel = void (*)(const RowParams);

// op: b200norm_op; tpr in {32,64,128,256,512}; ch in {1,2,3,4} or 0 (looping form, tpr 512 only).
// psmem (LayerNorm fwd only): weight / bias staged in shared memory as fp32.
RingKernel ring_fwd_kernel_f32(int op, int tpr, int ch, bool psmem);
RingKernel ring_fwd_kernel_bf16(int op, int tpr, int ch, bool psmem);
RingKernel ring_fwd_kernel_f16(int op, int tpr, int ch, bool psmem);
// forward norms with the fused residual add (RMSNORM_FWD / LAYERNORM_FWD only; ch in {1..4} or 0)
RingKernel ring_add_kernel_f32(int op, int tpr, int ch, bool psmem);
RingKernel ring_add_kernel_bf16(int op, int tpr, int ch, bool psmem);
RingKernel ring_add_kernel_f16(int op, int tpr, int ch, bool psmem);
// ... and with the fused dropout of x (benchmarks/python/test_dropout_*_fwd.py): ch in {1..4}, parameters in registers
RingKernel ring_drop_kernel_f32(int op, int tpr, int ch);
RingKernel ring_drop_kernel_bf16(int op, int tpr, int ch);
RingKernel ring_drop_kernel_f16(int op, int tpr, int ch);
// norm backward: cached (g / xc in registers, ch in {1,2}) or re-read (ch in {3,4}, tpr 512... any tpr).
RingKernel ring_bwd_kernel_f32(int op, int tpr, int ch, bool cached, bool centered);
RingKernel ring_bwd_kernel_bf16(int op, int tpr, int ch, bool cached, bool centered);
RingKernel ring_bwd_kernel_f16(int op, int tpr, int ch, bool cached, bool centered);
// fused-dropout backward (cached form): dy | x | residual | mask per stage, two gradient outputs
RingKernel ring_bwd_drop_kernel_f32(int op, int tpr, int ch);
RingKernel ring_bwd_drop_kernel_bf16(int op, int tpr, int ch);
RingKernel ring_bwd_drop_kernel_f16(int op, int tpr, int ch);
// instance norm backward (B200NORM_F_ROW_AFFINE)
RingKernel ring_bwd_rowaff_kernel_f32(int tpr, int ch, bool cached);
RingKernel ring_bwd_rowaff_kernel_bf16(int tpr, int ch, bool cached);
RingKernel ring_bwd_rowaff_kernel_f16(int tpr, int ch, bool cached);
// pipelined re-read backward (split-phase reductions overlapped across rows): ch in {3,4}, tpr in {64..512};
// lean: RMSNorm without the centring term on the instruction-trimmed kernel (tpr 256 / 512)
RingKernel ring_bwd_pipe_kernel_f32(int op, int tpr, int ch, bool lean);
RingKernel ring_bwd_pipe_kernel_bf16(int op, int tpr, int ch, bool lean);
RingKernel ring_bwd_pipe_kernel_f16(int op, int tpr, int ch, bool lean);
// pipelined RMSNorm forward (cluster rows)
RingKernel ring_fwd_pipe_kernel_f32(int op, int tpr, int ch);
RingKernel ring_fwd_pipe_kernel_bf16(int op, int tpr, int ch);
RingKernel ring_fwd_pipe_kernel_f16(int op, int tpr, int ch);

// pipelined softmax / log-softmax forward over cluster rows (kernels_softmax_pipe.cuh): tpr 512, ch in {2..8} or 0
RingKernel ring_sm_pipe_kernel_f32(int op, int tpr, int ch);
RingKernel ring_sm_pipe_kernel_bf16(int op, int tpr, int ch);
RingKernel ring_sm_pipe_kernel_f16(int op, int tpr, int ch);

// pipelined one-round LayerNorm forward (kernels_ln_pipe.cuh): tpr in {64..512}, ch in {1..4} (5, 6 at tpr 256), with or without the fused residual add
RingKernel ring_ln_pipe_kernel_f32(int tpr, int ch, bool add);
RingKernel ring_ln_pipe_kernel_bf16(int tpr, int ch, bool add);
RingKernel ring_ln_pipe_kernel_f16(int tpr, int ch, bool add);

// plain reductions on the ring (kernels_reduce.cuh): inner sum / max / fmax (ch 1..4 or 0 with tpr 512) and the
// column sum over the rows (ch 1..4)
RingKernel ring_reduce_inner_kernel_f32(int tpr, int ch);
RingKernel ring_reduce_inner_kernel_bf16(int tpr, int ch);
RingKernel ring_reduce_inner_kernel_f16(int tpr, int ch);
RingKernel ring_colsum_kernel_f32(int tpr, int ch);
RingKernel ring_colsum_kernel_bf16(int tpr, int ch);
RingKernel ring_colsum_kernel_f16(int tpr, int ch);

}  // namespace b200norm
