#!/bin/bash
# Round 2, visit 16 (one GPU): fused dropout on the ring kernels: parity, throughput.
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests/test_dropout_gpu.py tests/test_fused_add_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v16_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/v16_pytest.log
timeout 900 python scripts/kbench.py drop_ln_fwd:16384x4096:bf16 drop_rms_fwd:16384x4096:bf16 add_ln_fwd:16384x4096:bf16 add_rms_fwd:16384x4096:bf16 drop_ln_fwd:16384x8192:bf16 drop_rms_fwd:16384x8192:bf16 drop_ln_fwd:8192x12288:bf16 drop_ln_fwd:16384x4096:f32 drop_rms_fwd:16384x4096:f32 drop_ln_bwd:16384x4096:bf16 drop_rms_bwd:16384x4096:bf16 drop_ln_fwd:16384x4096:bf16:B200NORM_VARIANT=1 2>&1 | tee gpurun_out/v16_kbench.txt
