#!/bin/bash
# Round 2, visit 49 (one GPU): RMSNorm forward pipelined at five / six chunks per thread; LayerNorm / RMSNorm backward on the same rows.
set -u
mkdir -p gpurun_out
timeout 600 python scripts/kbench.py \
  rms_fwd:16384x24576:bf16 rms_fwd:16384x24576:bf16:B200NORM_PIPE=1 rms_fwd:16384x20480:bf16 rms_fwd:16384x20480:bf16:B200NORM_PIPE=1 \
  rms_fwd:16384x12288:f32 rms_fwd:16384x12288:f32:B200NORM_PIPE=1 \
  ln_bwd:16384x24576:bf16 rms_bwd:16384x24576:bf16 ln_bwd:16384x16384:bf16 ln_bwd:8192x32768:bf16 add_ln_fwd:16384x24576:bf16 add_rms_fwd:16384x24576:bf16 \
  2>&1 | tee gpurun_out/v49_kbench.txt
