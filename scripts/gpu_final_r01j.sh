#!/bin/bash
# Round-1 closing visit: full GPU parity suite, smoke, default bench + reference arm + the o1 workload,
# ncu launch list of the default bench, compute-sanitizer memcheck / racecheck over the outer-persistent tests.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench reference"; timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; head -c 600 gpurun_out/bench_reference.json; echo
echo "== bench default"; timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])
PY
: > gpurun_out/bench_all.jsonl
for w in c1 c3 c4 c5 o1; do
  echo "== bench $w"; timeout 600 python bench.py --workload $w --steps 100 --warmup 5 --no-e2e --no-cpu >> gpurun_out/bench_all.jsonl 2>> gpurun_out/bench_all.err; echo "rc=$?"
done
python - <<'PY'
import json
for l in open('gpurun_out/bench_all.jsonl'):
    d=json.loads(l); r=d['roofline']['per_kernel']
    print(d['config']['workload'][:48], 'value %.0f'%d['value'], '  '.join('%s %.1fus %.0fGB/s (%.1f%%)'%(k,v['ms']*1e3,v['gbs'],100*v['gbs']/d['roofline']['peak']) for k,v in r.items()))
PY
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== sanitizer (outer family)"
SEL='tests/test_outer_gpu.py::test_outer_cluster_split tests/test_outer_gpu.py::test_outer_edge_cases'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 5 python -m pytest $SEL -q -x -p no:cacheprovider -k "777x384-bf16 or edge" > gpurun_out/sanitize_outer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|Race reported|Invalid|hazard" gpurun_out/sanitize_outer_$tool.log | sort | uniq -c | head -8
done
