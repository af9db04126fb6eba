#!/bin/bash
# Round 2, visit 41 (one GPU): the pipelined softmax / cross-entropy kernel without a cluster for 48-75 KB rows (32 K vocabularies).
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_cross_entropy_gpu.py tests/test_vocab_gpu.py tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v41_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v41_pytest.log
timeout 900 python scripts/kbench.py \
  sm_fwd:8192x32064:bf16 sm_fwd:8192x32064:bf16:B200NORM_PIPE=-1 sm_fwd:8192x32064:bf16:B200NORM_NO_PIPE=1 \
  sm_fwd:16384x32768:bf16 sm_fwd:16384x32768:bf16:B200NORM_PIPE=-1 sm_fwd:16384x32768:bf16:B200NORM_NO_PIPE=1 \
  sm_fwd:16384x28672:bf16 sm_fwd:16384x28672:bf16:B200NORM_PIPE=-1 sm_fwd:8192x16384:f32 sm_fwd:8192x16384:f32:B200NORM_PIPE=-1 sm_fwd:8192x16384:f32:B200NORM_NO_PIPE=1 \
  ce_fwd:8192x32064:bf16 ce_fwd:8192x32064:bf16:B200NORM_PIPE=-1 ce_fwd:8192x32064:bf16:B200NORM_NO_PIPE=1 \
  ce_fwd:16384x32768:bf16 ce_fwd:16384x32768:bf16:B200NORM_PIPE=-1 ce_fwd:16384x32768:bf16:B200NORM_NO_PIPE=1 \
  sm_bwd:4096x131072:bf16 sm_bwd:8192x32064:bf16 sm_bwd:8192x65536:bf16 \
  2>&1 | tee gpurun_out/v41_kbench.txt
