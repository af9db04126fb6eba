#!/bin/bash
# Closing visit after the outer normalization work: full GPU parity suite, smoke, default bench, o1 bench,
# ncu full of the outer normalization kernels, sanitizer over the outer family.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-80
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench default"; timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"
echo "== bench o1"; timeout 600 python bench.py --workload o1 --steps 100 --warmup 5 > gpurun_out/bench_o1.json 2> gpurun_out/bench_o1.err; echo "rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_default.json','gpurun_out/bench_o1.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']['per_kernel']
    print(d['config']['workload'][:40], 'value %.0f'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.1f'%d['e2e']['value'], ' '.join('%s %.1fus %.1f%%'%(k,v['ms']*1e3,100*v['gbs']/d['roofline']['peak']) for k,v in r.items()), d['clocks'])
PY
echo "== ncu onorm"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:outer_softmax -s 8 -c 1 -o gpurun_out/prof_onorm_fwd -f python scripts/kbench.py onorm_fwd:8192x18432:bf16 > gpurun_out/ncu_onorm_fwd.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:outer_softmax -s 10 -c 1 -o gpurun_out/prof_onorm_bwd -f python scripts/kbench.py onorm_bwd:8192x18432:bf16 > gpurun_out/ncu_onorm_bwd.log 2>&1; echo "rc=$?"
echo "== sanitizer (outer family)"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 3 python -m pytest tests/test_outer_gpu.py -q -x -p no:cacheprovider -k "777x384-bf16 or edge or bwd_cluster or norm_cluster" > gpurun_out/sanitize_outer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|Race reported|Invalid|hazard" gpurun_out/sanitize_outer_$tool.log | cut -c1-160 | sort | uniq -c | head -6
done
