#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== outer tests"; timeout 900 python -m pytest tests/test_outer_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_outer.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_outer.log
timeout 900 python scripts/kbench.py onorm_bwd:128x262144:bf16 onorm_bwd:512x65536:bf16 onorm_bwd:2048x16384:bf16 onorm_bwd:8192x18432:bf16 onorm_bwd:16384x8192:bf16 \
   onorm_bwd:8192x4096:f32 onorm_fwd:128x262144:bf16 onorm_fwd:512x65536:bf16 onorm_fwd:2048x16384:bf16 onorm_fwd:8192x18432:bf16 onorm_fwd:25088x1024:bf16 onorm_bwd:25088x1024:bf16 2>&1 | tee gpurun_out/kbench_outer13.txt
