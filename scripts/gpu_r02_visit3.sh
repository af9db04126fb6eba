#!/bin/bash
# Round 2, visit 3 (one GPU): whole suite (new: reductions / var_mean / welford on the ring, thunder backward,
# LayerNorm pipe variants), knob sweeps for the LayerNorm forward, reduction kernels' bandwidth, ncu of the
# LayerNorm pipe kernel after the vectorised combine, timeline of the C4 / C2 backward.
set -u
mkdir -p gpurun_out
export B200NORM_PEER_TIMEOUT_S=20
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/v3_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/v3_pytest.log
echo "== kbench"
timeout 1200 python scripts/kbench.py ln_fwd:32768x12288:bf16 ln_fwd:32768x12288:bf16:B200NORM_DEPTH=5 ln_fwd:32768x12288:bf16:B200NORM_DEPTH=6 ln_fwd:32768x12288:bf16:B200NORM_DEPTH=8 \
  ln_fwd:32768x12288:bf16:B200NORM_TPR=256 ln_fwd:32768x12288:bf16:B200NORM_TPR=256,B200NORM_DEPTH=3 ln_fwd:32768x12288:bf16:B200NORM_DEPTH=3 \
  rms_fwd:32768x12288:bf16 rms_fwd:32768x12288:bf16:B200NORM_DEPTH=6 rms_fwd:32768x12288:bf16:B200NORM_DEPTH=8 \
  ln_fwd:16384x4096:bf16 ln_fwd:16384x4096:bf16:B200NORM_LN_PIPE=1 ln_fwd:16384x4096:bf16:B200NORM_LN_PIPE=1,B200NORM_DEPTH=6 ln_fwd:16384x4096:bf16:B200NORM_LN_PIPE=1,B200NORM_TPR=256 \
  ln_fwd:16384x4096:bf16:B200NORM_LN_PIPE=1,B200NORM_TPR=256,B200NORM_DEPTH=8 ln_fwd:16384x4096:bf16:B200NORM_TPR=256 ln_fwd:16384x4096:bf16:B200NORM_LN_PIPE=1,B200NORM_TPR=512 \
  ln_fwd:16384x8192:bf16 ln_fwd:16384x8192:bf16:B200NORM_LN_PIPE=1 ln_fwd:16384x8192:bf16:B200NORM_LN_PIPE=1,B200NORM_DEPTH=6 ln_fwd:16384x8192:bf16:B200NORM_LN_PIPE=1,B200NORM_TPR=512 ln_fwd:16384x8192:bf16:B200NORM_LN_PIPE=1,B200NORM_TPR=512,B200NORM_DEPTH=8 \
  ln_fwd:2048x4096:f32 ln_fwd:2048x4096:f32:B200NORM_LN_PIPE=1 ln_fwd:2048x4096:f32:B200NORM_LN_PIPE=1,B200NORM_TPR=512 \
  ln_fwd:4096x65536:bf16 ln_fwd:4096x65536:bf16:B200NORM_CLUSTER=8 ln_fwd:4096x65536:bf16:B200NORM_CLUSTER=4,B200NORM_NO_LN_PIPE=1 ln_fwd:16384x16384:bf16 ln_fwd:16384x16384:bf16:B200NORM_CLUSTER=2 \
  rms_bwd:4096x65536:bf16 rms_bwd:4096x65536:bf16:B200NORM_DEPTH=2 \
  red_sum:16384x4096:bf16 red_max:16384x4096:bf16 colsum:16384x4096:bf16 var_mean:16384x4096:bf16 red_sum:32768x12288:bf16 colsum:32768x12288:bf16 var_mean:32768x12288:bf16 \
  red_sum:16384x4096:f32 colsum:16384x4096:f32 var_mean:16384x4096:f32 red_sum:4096x65536:bf16 colsum:4096x65536:bf16 var_mean:4096x65536:bf16 2>&1 | tee gpurun_out/v3_kbench.txt
echo "== ncu"
bash scripts/ncu_box.sh ln_fwd_12288_b layernorm_fwd 4 ln_fwd:32768x12288:bf16
bash scripts/ncu_box.sh rms_fwd_12288 rmsnorm_fwd 4 rms_fwd:32768x12288:bf16
bash scripts/ncu_box.sh rms_fwd_4096 rmsnorm_fwd 4 rms_fwd:16384x4096:bf16
echo "== timelines"
timeout 600 python scripts/timeline.py ln_bwd:32768x12288:bf16 ln_bwd:4096x12288:bf16 rms_bwd:16384x4096:bf16 rms_bwd:4096x65536:bf16 ln_fwd:32768x12288:bf16 2>&1 | tee gpurun_out/v3_timeline.txt
du -sh gpurun_out
