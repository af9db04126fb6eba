#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== outer tests"; timeout 900 python -m pytest tests/test_outer_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_outer.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_outer.log
E=B200NORM_OUTER_NO_EARLY
timeout 900 python scripts/kbench.py \
  osm_fwd:128x262144:bf16 osm_fwd:128x262144:bf16:$E=1 osm_bwd:128x262144:bf16 osm_bwd:128x262144:bf16:$E=1 \
  osm_fwd:512x65536:bf16 osm_fwd:512x65536:bf16:$E=1 osm_bwd:512x65536:bf16 osm_bwd:512x65536:bf16:$E=1 \
  osm_fwd:2048x16384:bf16 osm_fwd:2048x16384:bf16:$E=1 osm_bwd:2048x16384:bf16 osm_bwd:2048x16384:bf16:$E=1 \
  osm_fwd:8192x18432:bf16 osm_fwd:8192x18432:bf16:$E=1 osm_bwd:8192x18432:bf16 osm_bwd:8192x18432:bf16:$E=1 \
  osm_fwd:16384x8192:bf16 osm_fwd:16384x8192:bf16:$E=1 osm_bwd:16384x8192:bf16 osm_bwd:16384x8192:bf16:$E=1 \
  osm_fwd:8192x4096:f32 osm_bwd:8192x4096:f32 osm_fwd:8192x4096:bf16 2>&1 | tee gpurun_out/kbench_outer8.txt
