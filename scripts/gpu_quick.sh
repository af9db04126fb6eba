#!/bin/bash
# Quick GPU visit: parity tests + per-config bench (+ optional env sweeps given as args "label:w:ENV=V,ENV=V").
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x --maxfail=5 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
b() {
  local label=$1; shift
  local w=$1; shift
  echo "== bench $w $label"; env "$@" timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/err.txt | tee gpurun_out/bench_${w}_${label}.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']['per_kernel']
    print('   value %.0f GB/s  ms/step %.4f  '%(d['value'],d['ms_per_step']) + '  '.join('%s %.1fus %.0fGB/s (%.1f%%)'%(k,v['ms']*1e3,v['gbs'],100*v['gbs']/6540.2) for k,v in r.items()))
"; tail -2 gpurun_out/err.txt
}
for w in c2 c3 c4 c5 c1; do b default $w X=1; done
for spec in "$@"; do
  IFS=: read -r label w envs <<< "$spec"
  b $label $w ${envs//,/ }
done
