#!/bin/bash
# Round 2, visit 33 (one GPU): producer path restored (extra streams compile-time): whole suite + every bench workload.
set -u
mkdir -p gpurun_out
export B200NORM_PEER_TIMEOUT_S=20
echo "== tests"; timeout 2400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/v33_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/v33_pytest.log
: > gpurun_out/v33_bench_all.jsonl
for w in c2 c1 c3 c4 c5 o1 o2 o3 r1; do timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | tail -1 >> gpurun_out/v33_bench_all.jsonl; done
python - <<'PY'
import json
for l in open('gpurun_out/v33_bench_all.jsonl'):
    if not l.startswith('{'): continue
    d=json.loads(l); print(d['config']['workload'][:46], d['validated'], round(d['value']), {k:(round(v['ms']*1e3,1), round(v['gbs'])) for k,v in d['roofline']['per_kernel'].items()}, round(d['roofline']['frac'],4))
PY
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/v33_bench_default.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/v33_bench_default.json').read().strip().splitlines()[-1]); print('default', d['validated'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'])"
timeout 600 python scripts/kbench.py rms_fwd:4096x65536:bf16 drop_ln_bwd:16384x4096:bf16 drop_ln_fwd:16384x4096:bf16 ce_bwd:4096x131072:bf16 onorm_fwd:100352x256:bf16 onorm_bwd:100352x256:bf16 ln_fwd:32768x12288:bf16 ln_fwd:16384x4096:bf16 add_ln_fwd:16384x4096:bf16 2>&1 | tee gpurun_out/v33_kbench.txt
