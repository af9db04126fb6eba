#!/bin/bash
# GPU visit 4: parity suite incl. fused-add + dynamic-schedule tests, timelines and kbench of the new defaults.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== timelines"
timeout 600 python scripts/timeline.py rms_fwd:16384x4096:bf16 rms_fwd:16384x4096:bf16:B200NORM_STATIC=1 rms_bwd:16384x4096:bf16 \
  rms_bwd:16384x4096:bf16:B200NORM_COOP=1 ln_bwd:32768x12288:bf16 sm_fwd:65536x8192:bf16 2>&1 | tee gpurun_out/timeline4.txt
echo "== kbench"
timeout 900 python scripts/kbench.py copy:16384x4096:bf16 rms_fwd:16384x4096:bf16 rms_fwd:16384x4096:bf16:B200NORM_STATIC=1 \
  rms_bwd:16384x4096:bf16 rms_bwd:16384x4096:bf16:B200NORM_COOP=1 add_rms_fwd:16384x4096:bf16 add_ln_fwd:16384x4096:bf16 \
  ln_fwd:16384x4096:bf16 ln_fwd:16384x4096:bf16:B200NORM_STATIC=1 ln_bwd:16384x4096:bf16 sm_fwd:16384x4096:bf16 sm_bwd:16384x4096:bf16 \
  sm_fwd:65536x8192:bf16 sm_fwd:65536x8192:bf16:B200NORM_STATIC=1 sm_bwd:65536x8192:bf16 sm_bwd:65536x8192:bf16:B200NORM_STATIC=1 \
  ln_bwd:32768x12288:bf16 ln_fwd:32768x12288:bf16 rms_fwd:32768x12288:bf16 rms_bwd:32768x12288:bf16 \
  rms_fwd:4096x65536:bf16 rms_fwd:4096x65536:bf16:B200NORM_CLUSTER=2 rms_bwd:4096x65536:bf16 ln_fwd:2048x4096:f32 \
  rms_fwd:65536x4096:bf16 rms_bwd:65536x4096:bf16 2>&1 | tee gpurun_out/kbench_v4.txt
for extra in "$@"; do
  echo "== extra: $extra"; timeout 900 bash -c "$extra"; echo "rc=$?"
done
