#!/bin/bash
# Round 2, visit 40 (one GPU): half of pass 1's exponentials on the FMA pipe (degree-5 polynomial), A/B through B200NORM_NO_POLY_EXP.
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_cross_entropy_gpu.py tests/test_vocab_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v40_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v40_pytest.log
timeout 900 python scripts/kbench.py \
  sm_fwd:4096x131072:bf16 sm_fwd:4096x131072:bf16:B200NORM_NO_POLY_EXP=1 sm_fwd:4096x100352:bf16 sm_fwd:4096x100352:bf16:B200NORM_NO_POLY_EXP=1 \
  sm_fwd:2048x256000:bf16 sm_fwd:2048x256000:bf16:B200NORM_NO_POLY_EXP=1 sm_fwd:2048x131072:f32 sm_fwd:2048x131072:f32:B200NORM_NO_POLY_EXP=1 \
  ce_fwd:4096x131072:bf16 ce_fwd:4096x131072:bf16:B200NORM_NO_POLY_EXP=1 ce_fwd:4096x100352:bf16 ce_fwd:4096x100352:bf16:B200NORM_NO_POLY_EXP=1 \
  ce_fwd:4096x151936:bf16 ce_fwd:4096x151936:bf16:B200NORM_NO_POLY_EXP=1 ce_fwd:2048x131072:f32 ce_fwd:2048x131072:f32:B200NORM_NO_POLY_EXP=1 \
  2>&1 | tee gpurun_out/v40_kbench.txt
