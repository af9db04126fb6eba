#!/bin/bash
# Round 2, visit 34 (one GPU): after removing the experimental softmax branch and making the instance-norm backward a compile-time flavour.
set -u
mkdir -p gpurun_out
export B200NORM_PEER_TIMEOUT_S=20
echo "== tests"; timeout 2400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/v34_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/v34_pytest.log
: > gpurun_out/v34_bench_all.jsonl
for w in c2 c3 c4 c5; do timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | tail -1 >> gpurun_out/v34_bench_all.jsonl; done
python - <<'PY'
import json
for l in open('gpurun_out/v34_bench_all.jsonl'):
    if not l.startswith('{'): continue
    d=json.loads(l); print(d['config']['workload'][:46], d['validated'], round(d['value']), {k:(round(v['ms']*1e3,1), round(v['gbs'])) for k,v in d['roofline']['per_kernel'].items()}, round(d['roofline']['frac'],4))
PY
timeout 600 python scripts/kbench.py drop_ln_bwd:16384x4096:bf16 drop_rms_bwd:16384x4096:bf16 inorm_bwd:32768x784:bf16 inorm_fwd:32768x784:bf16 sm_fwd:4096x131072:bf16 2>&1 | tee gpurun_out/v34_kbench.txt
