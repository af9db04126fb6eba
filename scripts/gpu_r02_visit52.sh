#!/bin/bash
# Round 2, visit 52 (one GPU): LayerNorm backward on wide rows — cached lock-step cluster plan against the pipelined re-read plan.
set -u
mkdir -p gpurun_out
timeout 300 python scripts/kbench.py \
  ln_bwd:16384x16384:bf16 ln_bwd:16384x16384:bf16:B200NORM_BWD_REREAD=1 ln_bwd:16384x16384:bf16:B200NORM_CLUSTER=4 \
  ln_bwd:8192x32768:bf16 ln_bwd:8192x32768:bf16:B200NORM_BWD_REREAD=1 ln_bwd:16384x20480:bf16 ln_bwd:16384x20480:bf16:B200NORM_BWD_REREAD=1 \
  ln_bwd:16384x8192:f32 ln_bwd:16384x8192:f32:B200NORM_BWD_REREAD=1 \
  2>&1 | tee gpurun_out/v52_kbench.txt
