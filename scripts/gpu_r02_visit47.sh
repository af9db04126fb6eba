#!/bin/bash
# Round 2, visit 47 (one GPU): ncu of the pipelined softmax without a cluster; where the softmax backward stands on one-group rows.
set -u
mkdir -p gpurun_out
bash scripts/ncu_box.sh r02u_sm_pipe_16384x32768 softmax_fwd_pipe 6 sm_fwd:16384x32768:bf16
timeout 600 python scripts/kbench.py \
  sm_bwd:16384x16384:bf16 sm_bwd:16384x20480:bf16 sm_bwd:16384x24576:bf16 sm_bwd:16384x32768:bf16 sm_bwd:16384x8192:f32 sm_bwd:8192x16384:f32 sm_bwd:32768x12288:bf16 \
  rms_fwd:16384x16384:bf16 rms_fwd:16384x24576:bf16 ln_fwd:16384x24576:bf16 rms_fwd:8192x32768:bf16 ln_fwd:8192x32768:bf16 \
  2>&1 | tee gpurun_out/v47_kbench.txt
