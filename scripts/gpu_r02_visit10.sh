#!/bin/bash
# Round 2, visit 10 (one GPU): ncu captures of the row-block outer normalization kernel.
set -u
mkdir -p gpurun_out
bash scripts/ncu_box.sh orows_fwd_100352x256 outer_rows_kernel 3 onorm_fwd:100352x256:bf16
bash scripts/ncu_box.sh orows_fwd_50000x64 outer_rows_kernel 3 onorm_fwd:50000x64:bf16 B200NORM_OUTER_ROWS=1
bash scripts/ncu_box.sh orows_bwd_100352x256 outer_rows_kernel 3 onorm_bwd:100352x256:bf16
