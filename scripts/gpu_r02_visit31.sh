#!/bin/bash
# Round 2, visit 31 (one GPU): two-row tiles for 8-12 KB inner reductions and the fused-add LayerNorm, finalize kernel for cluster backward plans.
set -u
mkdir -p gpurun_out
export B200NORM_PEER_TIMEOUT_S=20
echo "== tests"; timeout 2400 python -m pytest tests/test_fullsize_gpu.py tests/test_reduce_gpu.py tests/test_fused_add_gpu.py tests/test_parity_gpu.py tests/test_peer_gpu.py tests/test_frontend_gpu.py tests/test_dynamic_gpu.py tests/test_cross_entropy_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v31_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v31_pytest.log
timeout 900 python scripts/kbench.py red_sum:16384x4096:bf16 red_max:16384x4096:bf16 var_mean:16384x4096:bf16 red_sum:65536x4096:bf16 red_sum:16384x6144:bf16 red_sum:32768x12288:bf16 var_mean:32768x12288:bf16 ce_fwd:16384x4096:bf16 add_ln_fwd:16384x4096:bf16 add_ln_fwd:16384x4096:bf16:B200NORM_NO_TILE2=1 add_ln_fwd:65536x4096:bf16 add_ln_fwd:65536x4096:bf16:B200NORM_NO_TILE2=1 add_ln_fwd:16384x4096:f16 rms_bwd:4096x65536:bf16 rms_bwd:2048x131072:bf16 ln_bwd:2048x65536:bf16 2>&1 | tee gpurun_out/v31_kbench.txt
for w in c5 r1; do timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/v31_bench_$w.json 2>/dev/null; python -c "
import json,sys; d=json.loads(open('gpurun_out/v31_bench_$w.json').read().strip().splitlines()[-1]); print('$w', d['validated'], round(d['value']), {k:(round(v['ms']*1e3,1), round(v['gbs'])) for k,v in d['roofline']['per_kernel'].items()}, round(d['roofline']['frac'],4))"; done
