#!/bin/bash
# Round 2, visit 9 (one GPU): the row-block outer normalization kernel (outer_rows.cuh): parity, then throughput
# against the cluster-persistent family on batch-norm shapes.
set -u
mkdir -p gpurun_out
export B200NORM_PEER_TIMEOUT_S=20
echo "== tests"; timeout 1200 python -m pytest tests/test_outer_gpu.py tests/test_abi.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v9_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/v9_pytest.log
S="100352x256:bf16 401408x64:bf16 802816x64:bf16 200704x128:f32 1605632x64:bf16"
args=""
for s in $S; do for op in onorm_fwd onorm_bwd; do args="$args $op:$s"; done; done
S2="25088x1024:bf16 50176x512:bf16 12544x2048:bf16 6272x2048:bf16 8192x4096:bf16 16384x1024:f32 50000x64:bf16 40000x24:bf16 25088x256:bf16 12544x512:bf16 32768x128:bf16"
for s in $S2; do for op in onorm_fwd onorm_bwd; do args="$args $op:$s $op:$s:B200NORM_OUTER_ROWS=1"; done; done
timeout 900 python scripts/kbench.py $args 2>&1 | tee gpurun_out/v9_kbench.txt
python scripts/timeline.py onorm_fwd:100352x256:bf16 onorm_fwd:401408x64:bf16 2>&1 | tee gpurun_out/v9_timeline.txt
