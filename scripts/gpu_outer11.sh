#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== outer tests"; timeout 900 python -m pytest tests/test_outer_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_outer.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_outer.log
X=B200NORM_OUTER_TWO_EXCH
timeout 900 python scripts/kbench.py \
  osm_fwd:2048x16384:bf16 osm_fwd:2048x16384:bf16:$X=1 \
  osm_fwd:8192x18432:bf16 osm_fwd:8192x18432:bf16:$X=1 \
  osm_fwd:8192x4096:bf16 osm_fwd:8192x4096:bf16:$X=1 \
  osm_fwd:16384x8192:bf16 osm_fwd:16384x8192:bf16:$X=1 \
  osm_fwd:8192x4096:f32 osm_fwd:8192x4096:f32:$X=1 osm_fwd:512x4096:bf16 osm_fwd:512x4096:bf16:$X=1 2>&1 | tee gpurun_out/kbench_outer11.txt
