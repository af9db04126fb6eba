#!/bin/bash
# GPU visit 3: parity suite, timelines of the small-problem kernels, finish-variant A/B, ncu full of the
# kernels below 85 % (LayerNorm fwd, LayerNorm bwd C4, RMSNorm bwd C5).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== timelines"
timeout 600 python scripts/timeline.py rms_fwd:16384x4096:bf16 rms_bwd:16384x4096:bf16 rms_bwd:16384x4096:bf16:B200NORM_NO_COOP=1 \
  rms_bwd:16384x4096:bf16:B200NORM_NO_FUSED_FINISH=1 rms_bwd:65536x4096:bf16 ln_bwd:32768x12288:bf16 ln_fwd:16384x4096:bf16 \
  rms_bwd:4096x65536:bf16 ln_fwd:2048x4096:f32 2>&1 | tee gpurun_out/timeline.txt
echo "== kbench variants"
timeout 600 python scripts/kbench.py copy:16384x4096:bf16 rms_fwd:16384x4096:bf16 rms_bwd:16384x4096:bf16 rms_bwd:16384x4096:bf16:B200NORM_NO_COOP=1 \
  rms_bwd:16384x4096:bf16:B200NORM_NO_FUSED_FINISH=1 rms_bwd_dx:16384x4096:bf16 \
  ln_bwd:32768x12288:bf16 ln_bwd:32768x12288:bf16:B200NORM_NO_COOP=1 ln_bwd_dx:32768x12288:bf16 \
  rms_bwd:4096x65536:bf16 rms_bwd:4096x65536:bf16:B200NORM_NO_COOP=1 rms_bwd_dx:4096x65536:bf16 \
  ln_fwd:16384x4096:bf16 ln_fwd:32768x12288:bf16 ln_fwd:2048x4096:f32 2>&1 | tee gpurun_out/kbench_v3.txt
echo "== ncu full"
for spec in ln_fwd:16384x4096:bf16 ln_fwd:32768x12288:bf16 ln_bwd:32768x12288:bf16 rms_bwd:4096x65536:bf16; do
  tag=$(echo $spec | tr ':' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ring_kernel -s 4 -c 1 -o gpurun_out/prof_$tag -f python scripts/kbench.py $spec > gpurun_out/ncu_$tag.log 2>&1; echo "ncu $tag rc=$?"
done
for extra in "$@"; do
  echo "== extra: $extra"; timeout 900 bash -c "$extra"; echo "rc=$?"
done
ls -la gpurun_out | head -40
