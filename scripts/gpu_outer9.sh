#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== outer tests"; timeout 900 python -m pytest tests/test_outer_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_outer.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_outer.log
W=B200NORM_OUTER_W; C=B200NORM_CLUSTER
timeout 900 python scripts/kbench.py \
  osm_fwd:128x262144:bf16 osm_bwd:128x262144:bf16 \
  osm_fwd:512x65536:bf16 osm_fwd:512x65536:bf16:$W=256,$C=2 osm_fwd:512x65536:bf16:$W=512,$C=4 \
  osm_bwd:512x65536:bf16 osm_bwd:512x65536:bf16:$W=128,$C=2 osm_bwd:512x65536:bf16:$W=256,$C=4 \
  osm_fwd:2048x16384:bf16 osm_fwd:2048x16384:bf16:$W=128,$C=4 osm_fwd:2048x16384:bf16:$W=256,$C=8 osm_fwd:2048x16384:bf16:$W=512,$C=16 \
  osm_bwd:2048x16384:bf16 osm_bwd:2048x16384:bf16:$W=128,$C=8 osm_bwd:2048x16384:bf16:$W=256,$C=16 \
  osm_fwd:8192x18432:bf16 osm_fwd:8192x18432:bf16:$W=128,$C=16 \
  osm_bwd:8192x18432:bf16 \
  osm_fwd:16384x8192:bf16 osm_bwd:16384x8192:bf16 osm_fwd:8192x4096:f32 osm_bwd:8192x4096:f32 2>&1 | tee gpurun_out/kbench_outer9.txt
