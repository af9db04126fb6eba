#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
b() {
  local label=$1; shift
  local w=$1; shift
  echo "== bench $w $label"; env "$@" timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/err.txt | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']['per_kernel']
    print('   value %.0f GB/s  ms/step %.4f  '%(d['value'],d['ms_per_step']) + '  '.join('%s %.1fus %.0fGB/s (%.1f%%)'%(k,v['ms']*1e3,v['gbs'],100*v['gbs']/6547.8) for k,v in r.items()))
"; tail -2 gpurun_out/err.txt
}
b default c2 X=1
b depth3 c2 B200NORM_DEPTH=3
b tile2 c2 B200NORM_TILE_ROWS=2
b tile4 c2 B200NORM_TILE_ROWS=4
b tpr256 c2 B200NORM_TPR=256
b groups2 c2 B200NORM_GROUPS=2
b default c3 X=1
b default c4 X=1
b cluster1 c4 B200NORM_CLUSTER=1
b default c5 X=1
b cluster8 c5 B200NORM_CLUSTER=8
b cluster16 c5 B200NORM_CLUSTER=16
b default c1 X=1
echo "== ncu full c2"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tma_kernel -s 6 -c 2 -o gpurun_out/prof_c2_v3 -f python bench.py --workload c2 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
