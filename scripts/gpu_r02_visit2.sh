#!/bin/bash
# Round 2, visit 2 (one GPU): the pipelined one-round LayerNorm forward (tests + A/B), ncu captures reduced to
# text on the box, bench lines kept.
set -u
mkdir -p gpurun_out
echo "== tests (new)"; timeout 900 python -m pytest tests/test_fullsize_gpu.py tests/test_parity_gpu.py tests/test_fused_add_gpu.py tests/test_dynamic_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v2_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/v2_pytest.log
echo "== kbench LN fwd A/B"
timeout 900 python scripts/kbench.py ln_fwd:32768x12288:bf16 ln_fwd:32768x12288:bf16:B200NORM_NO_LN_PIPE=1 \
  ln_fwd:16384x4096:bf16 ln_fwd:16384x4096:bf16:B200NORM_LN_PIPE=1 ln_fwd:16384x8192:bf16 ln_fwd:16384x8192:bf16:B200NORM_LN_PIPE=1 \
  ln_fwd:65536x4096:bf16 ln_fwd:65536x4096:bf16:B200NORM_LN_PIPE=1 ln_fwd:2048x4096:f32 ln_fwd:2048x4096:f32:B200NORM_LN_PIPE=1 \
  ln_fwd:16384x4096:f32 ln_fwd:16384x4096:f32:B200NORM_LN_PIPE=1 ln_fwd:16384x16384:bf16 ln_fwd:16384x16384:bf16:B200NORM_NO_LN_PIPE=1 \
  ln_fwd:4096x65536:bf16 ln_fwd:4096x65536:bf16:B200NORM_NO_LN_PIPE=1 ln_fwd:16384x8192:f32 ln_fwd:16384x8192:f32:B200NORM_NO_LN_PIPE=1 \
  add_ln_fwd:16384x4096:bf16 add_ln_fwd:16384x12288:bf16 add_ln_fwd:16384x12288:bf16:B200NORM_NO_LN_PIPE=1 \
  ln_fwd:16384x2048:bf16 ln_fwd:32768x1024:bf16 ln_fwd:16384x5120:bf16 rms_fwd:16384x5120:bf16 2>&1 | tee gpurun_out/v2_kbench.txt
echo "== ncu"
bash scripts/ncu_box.sh c5_bwd norm_bwd_pipe 4 rms_bwd:4096x65536:bf16
bash scripts/ncu_box.sh ln_fwd_12288 layernorm_fwd 4 ln_fwd:32768x12288:bf16
bash scripts/ncu_box.sh ln_fwd_12288_old layernorm_fwd 4 ln_fwd:32768x12288:bf16 B200NORM_NO_LN_PIPE=1
bash scripts/ncu_box.sh ln_fwd_4096 layernorm_fwd 4 ln_fwd:16384x4096:bf16
bash scripts/ncu_box.sh ln_fwd_4096_pipe layernorm_fwd 4 ln_fwd:16384x4096:bf16 B200NORM_LN_PIPE=1
bash scripts/ncu_box.sh c3_fwd softmax_fwd 4 sm_fwd:65536x8192:bf16
bash scripts/ncu_box.sh c3_bwd softmax_bwd 4 sm_bwd:65536x8192:bf16
bash scripts/ncu_box.sh c2_bwd norm_bwd_ring 4 rms_bwd:16384x4096:bf16
echo "== bench default + launch list"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/v2_bench_default.json 2> gpurun_out/v2_bench_default.err; echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v2_launches_c2.csv python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-validate > gpurun_out/v2_launches.log 2>&1; echo "rc=$?"
du -sh gpurun_out
