#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== outer tests"; timeout 900 python -m pytest tests/test_outer_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_outer.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_outer.log
X=B200NORM_OUTER_NO_L2PF
specs=""
for s in osm_fwd:128x262144 osm_bwd:128x262144 osm_fwd:512x65536 osm_bwd:512x65536 osm_fwd:2048x16384 osm_bwd:2048x16384 osm_fwd:8192x18432 osm_bwd:8192x18432 osm_fwd:16384x8192 onorm_fwd:8192x18432 onorm_bwd:8192x18432 onorm_fwd:512x65536; do specs="$specs $s:bf16 $s:bf16:$X=1"; done
timeout 900 python scripts/kbench.py $specs osm_fwd:8192x4096:f32 osm_bwd:8192x4096:f32 2>&1 | tee gpurun_out/kbench_outer14.txt
