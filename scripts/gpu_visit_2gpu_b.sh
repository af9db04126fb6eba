#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== peer tests"; timeout 900 python -m pytest tests/test_peer_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -15
for w in c2 c4; do
  for coll in peer peer-fused nccl; do
  echo "== $w 2 GPUs $coll"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $w --collective $coll --steps 200 --warmup 10 --no-e2e --no-cpu 2> gpurun_out/err2.txt | tee gpurun_out/bench_${w}_n2_${coll}.json | python -c "
import sys,json
for l in sys.stdin:
    if not l.startswith('{'): continue
    d=json.loads(l); print('   value %.0f GB/s  ms/step %.4f launches %d scaling %s'%(d['value'],d['ms_per_step'],d['gpu_launches'],d['scaling']), {k:round(v['ms']*1e3,1) for k,v in d['roofline']['per_kernel'].items()}, d['config']['collective'][:40])"; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/err2.txt | tail -3
  done
done
