#!/bin/bash
# Round 2, visit 46 (one GPU): cross-entropy mode with 1024 threads per CTA (B200NORM_TPR=1024) — parity, then timing.
set -u
mkdir -p gpurun_out
echo "== tests (TPR=1024)"; B200NORM_TPR=1024 timeout 900 python -m pytest tests/test_cross_entropy_gpu.py -m gpu -q -x -p no:cacheprovider -k "pipe_many_rows or cluster_split or fwd_bwd" > gpurun_out/v46_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v46_pytest.log
timeout 600 python scripts/kbench.py \
  ce_fwd:16384x32768:bf16 ce_fwd:16384x32768:bf16:B200NORM_TPR=1024 ce_fwd:8192x32064:bf16 ce_fwd:8192x32064:bf16:B200NORM_TPR=1024 \
  ce_fwd:4096x131072:bf16 ce_fwd:4096x131072:bf16:B200NORM_TPR=1024 ce_fwd:4096x151936:bf16 ce_fwd:4096x151936:bf16:B200NORM_TPR=1024 \
  ce_fwd:2048x131072:f32 ce_fwd:2048x131072:f32:B200NORM_TPR=1024 ce_fwd:4096x100352:bf16 ce_fwd:4096x100352:bf16:B200NORM_TPR=1024 \
  ce_fwd:8192x16384:f32 ce_fwd:8192x16384:f32:B200NORM_TPR=1024 \
  2>&1 | tee gpurun_out/v46_kbench.txt
