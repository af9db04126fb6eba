#!/bin/bash
# Round 2, visit 12 (one GPU): row-block outer kernel with a plain launch + PDL; wide-row timeline.
set -u
mkdir -p gpurun_out
export B200NORM_PEER_TIMEOUT_S=20
echo "== tests"; timeout 1200 python -m pytest tests/test_outer_gpu.py tests/test_streams_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v12_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v12_pytest.log
S="100352x256:bf16 401408x64:bf16 1605632x64:bf16"
args=""
for s in $S; do for op in onorm_fwd onorm_bwd; do args="$args $op:$s $op:$s:B200NORM_COOP=1"; done; done
timeout 900 python scripts/kbench.py $args 2>&1 | tee gpurun_out/v12_kbench.txt
python scripts/timeline.py onorm_fwd:100352x256:bf16 onorm_bwd:100352x256:bf16 onorm_fwd:8192x4096:bf16:B200NORM_OUTER_ROWS=1 onorm_fwd:12544x2048:bf16:B200NORM_OUTER_ROWS=1 2>&1 | tee gpurun_out/v12_timeline.txt
