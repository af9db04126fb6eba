#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== outer tests"; timeout 900 python -m pytest tests/test_outer_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_outer.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_outer.log
timeout 600 python scripts/kbench.py onorm_fwd:100352x256:bf16 onorm_bwd:100352x256:bf16 onorm_fwd:401408x64:bf16 osm_fwd:262144x512:bf16 osm_fwd:2048x1000:f16 2>&1 | tee gpurun_out/kbench_outer15.txt
