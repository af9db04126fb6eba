#!/bin/bash
# GPU visit 6: pool-fraction sweep of the hybrid schedule + timelines of softmax bwd (static vs dynamic).
set -u
mkdir -p gpurun_out
echo "== timelines"
timeout 600 python scripts/timeline.py sm_bwd:65536x8192:bf16:B200NORM_STATIC=1 sm_bwd:65536x8192:bf16:B200NORM_POOL=1 \
  sm_fwd:65536x8192:bf16:B200NORM_STATIC=1 sm_fwd:65536x8192:bf16:B200NORM_POOL=1 2>&1 | tee gpurun_out/timeline6.txt
echo "== kbench"
specs=""
for op in rms_fwd:16384x4096 sm_fwd:65536x8192 sm_bwd:65536x8192 rms_fwd:65536x4096 ln_fwd:16384x4096 add_rms_fwd:16384x4096 sm_bwd:16384x4096 rms_fwd:32768x12288 ln_fwd:32768x12288; do
  for v in B200NORM_STATIC=1 B200NORM_POOL=1 B200NORM_POOL=2 B200NORM_POOL=4 B200NORM_POOL=8 B200NORM_STATIC=1; do specs="$specs $op:bf16:$v"; done
done
timeout 1200 python scripts/kbench.py $specs 2>&1 | tee gpurun_out/kbench_v6.txt
