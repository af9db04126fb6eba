#!/bin/bash
# Round 2, visit 18 (one GPU): dropout kernels as their own instantiations (plain fused-add kernels untouched): parity + throughput; o3 bench.
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests/test_dropout_gpu.py tests/test_fused_add_gpu.py tests/test_frontend_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v18_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v18_pytest.log
timeout 900 python scripts/kbench.py drop_ln_fwd:16384x4096:bf16 drop_rms_fwd:16384x4096:bf16 add_ln_fwd:16384x4096:bf16 add_rms_fwd:16384x4096:bf16 add_ln_fwd:16384x8192:bf16 drop_ln_fwd:16384x8192:bf16 drop_rms_fwd:16384x8192:bf16 drop_ln_fwd:16384x4096:f32 drop_rms_fwd:16384x4096:f32 drop_ln_bwd:16384x4096:bf16 drop_rms_bwd:16384x4096:bf16 drop_ln_bwd:16384x8192:bf16 drop_ln_bwd:16384x4096:f32 2>&1 | tee gpurun_out/v18_kbench.txt
timeout 600 python bench.py --workload o3 --steps 20 --warmup 5 --no-cpu > gpurun_out/v18_bench_o3.json 2>gpurun_out/v18_bench_o3.err; tail -c 1500 gpurun_out/v18_bench_o3.json
