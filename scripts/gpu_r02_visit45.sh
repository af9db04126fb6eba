#!/bin/bash
# Round 2, visit 45 (one GPU): cross-entropy forward on the multi-group reduction kernel over small clusters (several groups per CTA overlap each other) against the pipelined kernel.
set -u
mkdir -p gpurun_out
timeout 600 python scripts/kbench.py \
  ce_fwd:16384x32768:bf16 ce_fwd:16384x32768:bf16:B200NORM_NO_PIPE=1,B200NORM_CLUSTER=2 ce_fwd:16384x32768:bf16:B200NORM_NO_PIPE=1,B200NORM_CLUSTER=4 \
  ce_fwd:8192x32064:bf16:B200NORM_NO_PIPE=1,B200NORM_CLUSTER=2 \
  ce_fwd:4096x131072:bf16 ce_fwd:4096x131072:bf16:B200NORM_NO_PIPE=1,B200NORM_CLUSTER=8 ce_fwd:4096x131072:bf16:B200NORM_NO_PIPE=1,B200NORM_CLUSTER=16 \
  ce_fwd:4096x151936:bf16:B200NORM_NO_PIPE=1,B200NORM_CLUSTER=8 \
  2>&1 | tee gpurun_out/v45_kbench.txt
