#!/bin/bash
# Round 2, visit 20 (one GPU): fused cross-entropy (log-softmax + NLL) forward / backward: parity + throughput.
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests/test_cross_entropy_gpu.py tests/test_reduce_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v20_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/v20_pytest.log
timeout 900 python scripts/kbench.py ce_fwd:8192x32064:bf16 ce_bwd:8192x32064:bf16 ce_fwd:4096x131072:bf16 ce_bwd:4096x131072:bf16 ce_fwd:4096x151936:bf16 ce_bwd:4096x151936:bf16 ce_fwd:16384x32768:bf16 ce_bwd:16384x32768:bf16 ce_fwd:4096x131072:f32 ce_bwd:4096x131072:f32 red_max:4096x131072:bf16 sm_fwd:8192x32064:bf16 2>&1 | tee gpurun_out/v20_kbench.txt
