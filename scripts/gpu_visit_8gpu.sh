#!/bin/bash
# 8-GPU validation: fused peer reduce correctness at 8 ranks + C4 (strong) / C2 (weak) with both collectives.
set -u
mkdir -p gpurun_out
echo "== peer worker, 8 ranks"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/peer_worker.py fused 2>&1 | grep -v "OMP_NUM\|^\*\*\*" | tail -4
for spec in "c4 peer-fused" "c4 nccl" "c2 peer-fused" "c2 nccl"; do
  set -- $spec
  echo "== $1 8 GPUs $2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --workload $1 --collective $2 --steps 100 --warmup 5 --no-e2e --no-cpu 2> gpurun_out/err8.txt | tee gpurun_out/bench_$1_n8_$2.json | python -c "
import sys,json
for l in sys.stdin:
    if not l.startswith('{'): continue
    d=json.loads(l); print('   value %.0f GB/s  ms/step %.4f launches %d scaling %s'%(d['value'],d['ms_per_step'],d['gpu_launches'],d['scaling']), {k:round(v['ms']*1e3,1) for k,v in d['roofline']['per_kernel'].items()})"; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/err8.txt | tail -3
done
