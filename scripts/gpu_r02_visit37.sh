#!/bin/bash
# Round 2, visit 37 (one GPU): ncu of the pipelined cluster softmax / cross-entropy kernels; the slow vocabulary sizes.
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_cross_entropy_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v37_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v37_pytest.log
bash scripts/ncu_box.sh r02q_sm_pipe_4096x131072 softmax_fwd_pipe 6 sm_fwd:4096x131072:bf16
bash scripts/ncu_box.sh r02q_ce_pipe_4096x131072 softmax_fwd_pipe 6 ce_fwd:4096x131072:bf16
timeout 900 python scripts/kbench.py \
  sm_fwd:4096x100352:bf16 sm_fwd:4096x100352:bf16:B200NORM_DEPTH=3 sm_fwd:4096x100352:bf16:B200NORM_NO_PIPE=1 sm_fwd:4096x100352:bf16:B200NORM_CLUSTER=8 sm_fwd:4096x100352:bf16:B200NORM_CLUSTER=2 \
  sm_fwd:4096x98304:bf16 sm_fwd:4096x114688:bf16 \
  sm_fwd:2048x202048:bf16 sm_fwd:2048x202048:bf16:B200NORM_DEPTH=3 sm_fwd:2048x202048:bf16:B200NORM_CLUSTER=16 sm_fwd:2048x262144:bf16 \
  2>&1 | tee gpurun_out/v37_kbench.txt
