#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== outer tests"; timeout 900 python -m pytest tests/test_outer_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_outer.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_outer.log
W=B200NORM_OUTER_W; C=B200NORM_CLUSTER; D=B200NORM_OUTER_DEPTH; P=B200NORM_OUTER_PER_SM; S=B200NORM_OUTER_SKIP_REDUCE
timeout 900 python scripts/kbench.py \
  osm_fwd:128x262144:bf16 osm_fwd:128x262144:bf16:$P=4 osm_fwd:128x262144:bf16:$P=2 osm_fwd:128x262144:bf16:$S=1 osm_fwd:128x262144:bf16:$P=4,$S=1 \
  osm_fwd:512x65536:bf16 osm_fwd:512x65536:bf16:$P=4 osm_fwd:512x65536:bf16:$S=1 \
  osm_fwd:2048x16384:bf16 osm_fwd:2048x16384:bf16:$P=4 osm_fwd:2048x16384:bf16:$S=1 \
  osm_fwd:8192x18432:bf16 osm_fwd:8192x18432:bf16:$P=4 osm_fwd:8192x18432:bf16:$S=1 \
  osm_fwd:16384x8192:bf16 osm_fwd:16384x8192:bf16:$P=4 \
  osm_bwd:512x65536:bf16 osm_bwd:512x65536:bf16:$P=4 osm_bwd:512x65536:bf16:$S=1 \
  osm_bwd:8192x18432:bf16 osm_bwd:8192x18432:bf16:$P=4 osm_bwd:8192x18432:bf16:$S=1 \
  osm_fwd:8192x4096:f32 osm_fwd:8192x4096:f32:$P=4 2>&1 | tee gpurun_out/kbench_outer5.txt
