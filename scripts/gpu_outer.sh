#!/bin/bash
# GPU visit: parity of the outer-persistent softmax family + slab-width / depth A/B at the reference's
# outer_persistent benchmark shapes (benchmarks/python/test_softmax_{fwd,bwd}.py, reduction_axis = 0).
set -u
mkdir -p gpurun_out
echo "== outer tests"; timeout 900 python -m pytest tests/test_outer_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_outer.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_outer.log
echo "== kbench outer"
timeout 900 python scripts/kbench.py \
  osm_fwd:512x16384:bf16 osm_fwd:512x16384:bf16:B200NORM_OUTER_W=256 osm_fwd:512x16384:bf16:B200NORM_OUTER_W=512 \
  osm_fwd:2048x16384:bf16 osm_fwd:2048x16384:bf16:B200NORM_OUTER_W=256 osm_fwd:2048x16384:bf16:B200NORM_OUTER_W=64 \
  osm_fwd:8192x4096:bf16 osm_fwd:8192x4096:bf16:B200NORM_OUTER_W=128 osm_fwd:8192x4096:bf16:B200NORM_OUTER_W=128,B200NORM_OUTER_DEPTH=1 \
  osm_fwd:8192x18432:bf16 osm_fwd:8192x18432:bf16:B200NORM_OUTER_W=128 \
  osm_fwd:16384x8192:bf16 osm_fwd:16384x8192:bf16:B200NORM_OUTER_W=128 \
  osm_bwd:2048x16384:bf16 osm_bwd:2048x16384:bf16:B200NORM_OUTER_W=64 osm_bwd:8192x4096:bf16 osm_bwd:8192x4096:bf16:B200NORM_OUTER_W=128 \
  osm_bwd:8192x18432:bf16 osm_bwd:16384x8192:bf16 \
  osm_fwd:8192x4096:f32 osm_fwd:2048x16384:f32 osm_bwd:8192x4096:f32 osm_fwd:16x73728:bf16 osm_fwd:512x4096:bf16 \
  sm_fwd:16384x8192:bf16 copy:8192x18432:bf16 2>&1 | tee gpurun_out/kbench_outer.txt
for extra in "$@"; do
  echo "== extra: $extra"; timeout 900 bash -c "$extra"; echo "rc=$?"
done
