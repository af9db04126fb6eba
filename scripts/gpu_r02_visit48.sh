#!/bin/bash
# Round 2, visit 48 (one GPU): LayerNorm forward with 5 / 6 chunks per thread at 512 threads (no cluster for N = 20480 / 24576).
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_fullsize_gpu.py tests/test_parity_gpu.py tests/test_instance_norm_gpu.py tests/test_reduce_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v48_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v48_pytest.log
timeout 600 python scripts/kbench.py \
  ln_fwd:16384x24576:bf16 ln_fwd:16384x24576:bf16:B200NORM_LN_PIPE=-1 ln_fwd:16384x20480:bf16 ln_fwd:16384x20480:bf16:B200NORM_LN_PIPE=-1 \
  ln_fwd:16384x12288:f32 ln_fwd:16384x12288:f32:B200NORM_LN_PIPE=-1 ln_fwd:16384x10240:f32 ln_fwd:16384x10240:f32:B200NORM_LN_PIPE=-1 \
  ln_fwd:8192x32768:bf16 ln_fwd:16384x16384:bf16 \
  2>&1 | tee gpurun_out/v48_kbench.txt
