#!/bin/bash
# Round 2, visit 36 (one GPU): cross-entropy flavour of the pipelined cluster softmax; smallest-cluster rule; fast exp in the fp32 sum.
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_vocab_gpu.py tests/test_cross_entropy_gpu.py tests/test_reduce_gpu.py tests/test_abi.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v36_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/v36_pytest.log
timeout 900 python scripts/kbench.py \
  ce_fwd:4096x131072:bf16 ce_fwd:4096x131072:bf16:B200NORM_NO_PIPE=1 ce_fwd:4096x131072:bf16:B200NORM_CLUSTER=8 ce_fwd:4096x131072:bf16:B200NORM_DEPTH=2 \
  ce_fwd:4096x151936:bf16 ce_fwd:4096x151936:bf16:B200NORM_NO_PIPE=1 ce_fwd:8192x32064:bf16 ce_fwd:16384x32768:bf16 ce_fwd:16384x32768:bf16:B200NORM_CLUSTER=2 \
  ce_fwd:2048x131072:f32 ce_fwd:2048x131072:f32:B200NORM_NO_PIPE=1 \
  sm_fwd:4096x152064:bf16 sm_fwd:4096x152064:bf16:B200NORM_CLUSTER=8 sm_fwd:2048x131072:f32 sm_fwd:4096x131072:bf16 sm_fwd:4096x100352:bf16 sm_fwd:4096x129280:bf16 sm_fwd:2048x202048:bf16 \
  2>&1 | tee gpurun_out/v36_kbench.txt
