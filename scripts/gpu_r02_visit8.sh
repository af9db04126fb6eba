#!/bin/bash
# Round 2, visit 8 (one GPU): the CEN = false instantiations of the ring backward (textbook RMSNorm): parity + A/B.
set -u
mkdir -p gpurun_out
export B200NORM_PEER_TIMEOUT_S=20
echo "== tests"; timeout 900 python -m pytest tests/test_fullsize_gpu.py tests/test_parity_gpu.py tests/test_peer_gpu.py tests/test_frontend_gpu.py tests/test_dynamic_gpu.py tests/test_streams_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v8_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v8_pytest.log
timeout 900 python scripts/kbench.py rms_bwd:16384x4096:bf16 rms_bwd:16384x4096:bf16:B200NORM_NO_LEAN=1 rms_bwd:16384x4096:bf16 rms_bwd:16384x4096:bf16:B200NORM_NO_LEAN=1 rms_bwd:65536x4096:bf16 rms_bwd:65536x4096:bf16:B200NORM_NO_LEAN=1 \
  rms_bwd:16384x8192:bf16 rms_bwd:16384x8192:bf16:B200NORM_NO_LEAN=1 rms_bwd:32768x12288:bf16 rms_bwd:32768x12288:bf16:B200NORM_NO_LEAN=1 rms_bwd:16384x4096:f32 rms_bwd:16384x4096:f32:B200NORM_NO_LEAN=1 2>&1 | tee gpurun_out/v8_kbench.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/v8_bench_default.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/v8_bench_default.json').read().strip().splitlines()[-1]); print(d['validated'], d['value'], {k:(round(v['ms']*1e3,1), round(v['gbs'])) for k,v in d['roofline']['per_kernel'].items()}, d['roofline']['frac'], d['e2e']['value'])"
