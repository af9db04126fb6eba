#!/bin/bash
# Round 2, visit 22 (one GPU): instance norm (row-affine LayerNorm kernels): parity + throughput.
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests/test_instance_norm_gpu.py tests/test_parity_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v22_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/v22_pytest.log
timeout 900 python scripts/kbench.py inorm_fwd:32768x784:bf16 inorm_bwd:32768x784:bf16 inorm_fwd:4096x1024:f32 inorm_bwd:4096x1024:f32 inorm_fwd:32768x3136:bf16 inorm_bwd:32768x3136:bf16 inorm_fwd:8192x12544:bf16 inorm_bwd:8192x12544:bf16 ln_fwd:32768x784:bf16 ln_bwd:32768x784:bf16 ln_fwd:32768x3136:bf16 ln_bwd:32768x3136:bf16 2>&1 | tee gpurun_out/v22_kbench.txt
