#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== outer tests"; timeout 900 python -m pytest tests/test_outer_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_outer.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_outer.log
echo "== kbench outer"
timeout 900 python scripts/kbench.py \
  osm_fwd:128x262144:bf16 osm_fwd:128x262144:bf16:B200NORM_OUTER_W=128 osm_fwd:128x262144:bf16:B200NORM_OUTER_W=512 \
  osm_fwd:512x65536:bf16 osm_fwd:512x65536:bf16:B200NORM_CLUSTER=1 osm_fwd:512x65536:bf16:B200NORM_CLUSTER=4 osm_fwd:512x65536:bf16:B200NORM_CLUSTER=16 \
  osm_fwd:512x65536:bf16:B200NORM_CLUSTER=2,B200NORM_OUTER_W=256 \
  osm_fwd:2048x16384:bf16 osm_fwd:2048x16384:bf16:B200NORM_OUTER_W=64 osm_fwd:2048x16384:bf16:B200NORM_CLUSTER=4 osm_fwd:2048x16384:bf16:B200NORM_CLUSTER=16 \
  osm_fwd:8192x4096:bf16 osm_fwd:8192x4096:bf16:B200NORM_OUTER_W=128 \
  osm_fwd:8192x18432:bf16 osm_fwd:8192x18432:bf16:B200NORM_OUTER_W=128 \
  osm_fwd:16384x8192:bf16 osm_fwd:16384x8192:bf16:B200NORM_OUTER_W=128 \
  osm_bwd:512x65536:bf16 osm_bwd:2048x16384:bf16 osm_bwd:2048x16384:bf16:B200NORM_OUTER_W=64 osm_bwd:8192x4096:bf16 osm_bwd:8192x4096:bf16:B200NORM_OUTER_W=128 \
  osm_bwd:8192x18432:bf16 osm_bwd:16384x8192:bf16 \
  osm_fwd:8192x4096:f32 osm_bwd:8192x4096:f32 osm_fwd:2048x16384:f32 2>&1 | tee gpurun_out/kbench_outer3.txt
echo "== ncu"
for spec in osm_fwd:8192x18432:bf16; do
  tag=$(echo $spec | tr ':' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:outer_softmax -s 4 -c 1 -o gpurun_out/prof3_$tag -f python scripts/kbench.py $spec > gpurun_out/ncu_$tag.log 2>&1; echo "ncu $tag rc=$?"
done
