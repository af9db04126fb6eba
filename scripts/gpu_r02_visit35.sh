#!/bin/bash
# Round 2, visit 35 (one GPU): pipelined one-exchange cluster softmax forward — parity, then timing against the register kernel.
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_vocab_gpu.py tests/test_parity_gpu.py tests/test_abi.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v35_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/v35_pytest.log
timeout 900 python scripts/kbench.py \
  sm_fwd:4096x131072:bf16 sm_fwd:4096x131072:bf16:B200NORM_CLUSTER=8 sm_fwd:4096x131072:bf16:B200NORM_CLUSTER=2 sm_fwd:4096x131072:bf16:B200NORM_NO_PIPE=1 \
  sm_fwd:4096x131072:bf16:B200NORM_DEPTH=2 sm_fwd:4096x131072:bf16:B200NORM_CLUSTER=8,B200NORM_DEPTH=3 \
  sm_fwd:2048x256000:bf16 sm_fwd:2048x256000:bf16:B200NORM_CLUSTER=16 sm_fwd:2048x256000:bf16:B200NORM_NO_PIPE=1 \
  sm_fwd:8192x65536:bf16 sm_fwd:8192x65536:bf16:B200NORM_CLUSTER=4 sm_fwd:8192x65536:bf16:B200NORM_NO_PIPE=1 \
  sm_fwd:2048x131072:f32 sm_fwd:2048x131072:f32:B200NORM_NO_PIPE=1 sm_fwd:4096x152064:bf16 sm_fwd:4096x152064:bf16:B200NORM_NO_PIPE=1 \
  2>&1 | tee gpurun_out/v35_kbench.txt
