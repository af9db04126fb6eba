#!/bin/bash
# Round 2, visit 43 (one GPU): cross-entropy mode one step further apart (merge a pass 1 after the post, collect an iteration after the send); planner rule for one-group softmax rows.
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_cross_entropy_gpu.py tests/test_vocab_gpu.py tests/test_frontend_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v43_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v43_pytest.log
timeout 900 python scripts/kbench.py \
  ce_fwd:8192x32064:bf16 ce_fwd:16384x32768:bf16 ce_fwd:4096x131072:bf16 ce_fwd:4096x151936:bf16 ce_fwd:4096x100352:bf16 ce_fwd:2048x131072:f32 ce_fwd:8192x65536:bf16 ce_fwd:2048x256000:bf16 \
  sm_fwd:16384x24576:bf16 sm_fwd:16384x16384:bf16 sm_fwd:16384x8192:f32 \
  2>&1 | tee gpurun_out/v43_kbench.txt
