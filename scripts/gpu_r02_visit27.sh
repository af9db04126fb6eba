#!/bin/bash
# Round 2, visit 27 (one GPU): sixteen one-warp groups for short forward rows: the whole suite + throughput.
set -u
mkdir -p gpurun_out
export B200NORM_PEER_TIMEOUT_S=20
echo "== tests"; timeout 2400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/v27_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v27_pytest.log
args=""
for s in 32768x784:bf16 65536x1024:bf16 131072x512:bf16 262144x256:bf16 32768x1000:f32 65536x512:f32; do for op in rms_fwd ln_fwd sm_fwd sm_bwd add_rms_fwd; do args="$args $op:$s $op:$s:B200NORM_GROUPS=8"; done; done
timeout 900 python scripts/kbench.py $args inorm_fwd:32768x784:bf16 drop_ln_fwd:65536x1024:bf16 drop_ln_fwd:65536x1024:bf16:B200NORM_GROUPS=8 2>&1 | tee gpurun_out/v27_kbench.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/v27_bench_default.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/v27_bench_default.json').read().strip().splitlines()[-1]); print(d['validated'], d['value'], {k:(round(v['ms']*1e3,1), round(v['gbs'])) for k,v in d['roofline']['per_kernel'].items()}, d['roofline']['frac'], d['e2e']['value'])"
timeout 600 python bench.py --workload c1 --steps 50 --warmup 5 --no-cpu --no-e2e > gpurun_out/v27_bench_c1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/v27_bench_c1.json').read().strip().splitlines()[-1]); print('c1', d['validated'], d['value'], d['roofline']['frac'])"
