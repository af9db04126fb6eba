#!/bin/bash
# Round 2, visit 42 (one GPU): where does the pipelined softmax beat the register kernel without a cluster? (B200NORM_PIPE=1 forces it)
set -u
mkdir -p gpurun_out
timeout 900 python scripts/kbench.py \
  sm_fwd:16384x16384:bf16 sm_fwd:16384x16384:bf16:B200NORM_PIPE=1 sm_fwd:16384x24576:bf16 sm_fwd:16384x24576:bf16:B200NORM_PIPE=1 \
  sm_fwd:16384x20480:bf16 sm_fwd:16384x20480:bf16:B200NORM_PIPE=1 sm_fwd:32768x12288:bf16 sm_fwd:32768x12288:bf16:B200NORM_PIPE=1 \
  sm_fwd:65536x8192:bf16 sm_fwd:65536x8192:bf16:B200NORM_PIPE=1 sm_fwd:16384x8192:f32 sm_fwd:16384x8192:f32:B200NORM_PIPE=1 sm_fwd:8192x12288:f32 sm_fwd:8192x12288:f32:B200NORM_PIPE=1 \
  ce_fwd:16384x16384:bf16 ce_fwd:16384x16384:bf16:B200NORM_PIPE=1 ce_fwd:16384x24576:bf16 ce_fwd:16384x24576:bf16:B200NORM_PIPE=1 \
  2>&1 | tee gpurun_out/v42_kbench.txt
