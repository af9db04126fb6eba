#!/bin/bash
# Round 2, visit 23 (one GPU): dropout forward with <= 2 chunks per thread.
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests/test_dropout_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v23_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/v23_pytest.log
timeout 900 python scripts/kbench.py drop_ln_fwd:16384x4096:bf16 drop_ln_fwd:16384x4096:bf16:B200NORM_DROP_CH4=1 drop_rms_fwd:16384x4096:bf16 drop_rms_fwd:16384x4096:bf16:B200NORM_DROP_CH4=1 drop_ln_fwd:16384x8192:bf16 drop_ln_fwd:16384x8192:bf16:B200NORM_DROP_CH4=1 drop_rms_fwd:16384x8192:bf16 drop_rms_fwd:16384x8192:bf16:B200NORM_DROP_CH4=1 drop_ln_fwd:16384x4096:f32 drop_ln_fwd:16384x4096:f32:B200NORM_DROP_CH4=1 drop_ln_fwd:32768x2048:bf16 drop_ln_fwd:32768x2048:bf16:B200NORM_DROP_CH4=1 2>&1 | tee gpurun_out/v23_kbench.txt
