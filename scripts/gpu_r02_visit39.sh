#!/bin/bash
# Round 2, visit 39 (one GPU): balanced tail (4-byte pieces) of the pipelined cluster softmax / cross-entropy.
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_cross_entropy_gpu.py tests/test_vocab_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v39_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v39_pytest.log
timeout 900 python scripts/kbench.py \
  sm_fwd:4096x100352:bf16 sm_fwd:4096x100352:bf16:B200NORM_NO_TILE2=1 sm_fwd:2048x202048:bf16 sm_fwd:2048x202048:bf16:B200NORM_NO_TILE2=1 \
  sm_fwd:4096x131072:bf16 sm_fwd:4096x133120:bf16 sm_fwd:4096x133120:bf16:B200NORM_NO_TILE2=1 sm_fwd:2048x100352:f32 sm_fwd:2048x100352:f32:B200NORM_NO_TILE2=1 \
  ce_fwd:4096x100352:bf16 ce_fwd:4096x100352:bf16:B200NORM_NO_TILE2=1 ce_fwd:4096x131072:bf16 \
  2>&1 | tee gpurun_out/v39_kbench.txt
