#!/bin/bash
# Round 2, visit 38 (one GPU): cross-entropy mode with only the first warp reducing; ncu of the slow N = 100352 softmax.
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_cross_entropy_gpu.py tests/test_vocab_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v38_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v38_pytest.log
timeout 900 python scripts/kbench.py \
  ce_fwd:4096x131072:bf16 ce_fwd:4096x131072:bf16:B200NORM_DEPTH=2 ce_fwd:4096x151936:bf16 ce_fwd:2048x131072:f32 ce_fwd:8192x65536:bf16 ce_fwd:2048x256000:bf16 ce_fwd:4096x100352:bf16 \
  2>&1 | tee gpurun_out/v38_kbench.txt
bash scripts/ncu_box.sh r02r_sm_pipe_4096x100352 softmax_fwd_pipe 6 sm_fwd:4096x100352:bf16
bash scripts/ncu_box.sh r02r_sm_pipe_4096x98304 softmax_fwd_pipe 6 sm_fwd:4096x98304:bf16
