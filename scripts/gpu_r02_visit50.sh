#!/bin/bash
# Round 2, visit 50 (one GPU): full GPU suite, smoke, every bench workload (validated), default bench with e2e / cpu legs, reference arm, launch list.
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 2400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/v50_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/v50_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
: > gpurun_out/v50_bench_all.jsonl
for w in c1 c2 c3 c4 c5 o1 o2 o3 r1 v1 v2; do timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | tail -1 >> gpurun_out/v50_bench_all.jsonl; done
python - <<'PY'
import json
for l in open('gpurun_out/v50_bench_all.jsonl'):
    if not l.startswith('{'): continue
    d=json.loads(l); print(d['config']['workload'][:60], d.get('validated'), round(d['value']), {k:(round(v['ms']*1e3,1), round(v['gbs'])) for k,v in d['roofline']['per_kernel'].items()}, round(d['roofline']['frac'],4))
PY
echo "== default bench"; timeout 900 python bench.py > gpurun_out/v50_bench_default.json 2>gpurun_out/v50_bench_default.err; echo "rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/v50_bench_default.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 | tail -1 | cut -c1-400
