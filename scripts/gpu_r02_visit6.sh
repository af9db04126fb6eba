#!/bin/bash
# Round 2, visit 6 (one GPU): the lean RMSNorm backward pipe kernel (C5): parity tests, A/B, ncu.
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_fullsize_gpu.py tests/test_parity_gpu.py tests/test_peer_gpu.py tests/test_frontend_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v6_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/v6_pytest.log
echo "== kbench"
timeout 900 python scripts/kbench.py rms_bwd:4096x65536:bf16 rms_bwd:4096x65536:bf16:B200NORM_NO_LEAN=1 rms_bwd_dx:4096x65536:bf16 rms_bwd:4096x65536:bf16 rms_bwd:4096x65536:bf16:B200NORM_NO_LEAN=1 \
  rms_bwd:2048x131072:bf16 rms_bwd:2048x131072:bf16:B200NORM_NO_LEAN=1 rms_bwd:16384x8192:f32 rms_bwd:16384x8192:f32:B200NORM_NO_LEAN=1 rms_bwd:8192x32768:bf16 rms_bwd:8192x32768:bf16:B200NORM_NO_LEAN=1 \
  rms_bwd:32768x12288:bf16 rms_bwd:32768x12288:bf16:B200NORM_PIPE=1 rms_bwd:32768x12288:bf16:B200NORM_PIPE=1,B200NORM_NO_LEAN=1 2>&1 | tee gpurun_out/v6_kbench.txt
bash scripts/ncu_box.sh c5_bwd_lean rms_bwd_pipe_lean 4 rms_bwd:4096x65536:bf16
timeout 600 python bench.py --workload c5 --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/v6_bench_c5.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/v6_bench_c5.json').read().strip().splitlines()[-1]); print(d['validated'], d['value'], {k:(round(v['ms']*1e3,1), round(v['gbs'])) for k,v in d['roofline']['per_kernel'].items()}, d['roofline']['frac'])"
