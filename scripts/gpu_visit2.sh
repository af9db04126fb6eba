#!/bin/bash
# One GPU-box visit: smoke, parity tests, bench of every config, full-contract bench, reference arm,
# ncu launch list and full captures of the C2 forward / backward kernels. Every step is under `timeout`.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
b() {
  local label=$1; shift
  local w=$1; shift
  echo "== bench $w $label"; env "$@" timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/err.txt | tee gpurun_out/bench_${w}_${label}.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']['per_kernel']; pk=d['roofline']['peak']
    print('   value %.0f GB/s  ms/step %.4f  '%(d['value'],d['ms_per_step']) + '  '.join('%s %.1fus %.0fGB/s (%.1f%%)'%(k,v['ms']*1e3,v['gbs'],100*v['gbs']/pk) for k,v in r.items()))
"; tail -2 gpurun_out/err.txt
}
for w in c2 c3 c4 c5 c1; do b default $w X=1; done
echo "== bench default (full contract)"; timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; cat gpurun_out/bench_reference.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload c2 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full c2 (fwd + bwd)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ring_kernel -s 6 -c 2 -o gpurun_out/prof_c2 -f python bench.py --workload c2 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
for extra in "$@"; do
  echo "== extra: $extra"; timeout 900 bash -c "$extra"; echo "rc=$?"
done
ls -la gpurun_out
