#!/bin/bash
# Round 2, visit 5 (one GPU): whole suite (var_mean on the multi-group reduction kernel, vocabulary-size rows,
# residual-add LayerNorm pipe at every width), kernel numbers of the changed paths.
set -u
mkdir -p gpurun_out
export B200NORM_PEER_TIMEOUT_S=20
echo "== tests"; timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/v5_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/v5_pytest.log
echo "== kbench"
timeout 1200 python scripts/kbench.py red_sum:16384x4096:bf16 red_max:16384x4096:bf16 var_mean:16384x4096:bf16 red_sum:32768x12288:bf16 red_max:32768x12288:bf16 var_mean:32768x12288:bf16 \
  red_sum:65536x8192:bf16 var_mean:65536x8192:bf16 var_mean:16384x4096:f32 var_mean:4096x65536:bf16 red_sum:4096x65536:bf16 \
  add_ln_fwd:16384x4096:bf16 add_ln_fwd:16384x8192:bf16 add_rms_fwd:16384x4096:bf16 sm_fwd:65536x8192:bf16 sm_fwd:18x1048576:bf16 sm_fwd:4096x131072:bf16 sm_fwd:2048x256000:bf16 \
  ln_fwd:32768x12288:bf16 ln_fwd:16384x4096:bf16 2>&1 | tee gpurun_out/v5_kbench.txt
du -sh gpurun_out
