#!/bin/bash
# 2-GPU visit: the row-sharded paths over NCCL (C2 weak scaling, C4 strong scaling), 1-GPU lines beside them.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus2.txt 2>&1
for w in c2 c4; do
  echo "== $w 1 GPU"; timeout 300 python bench.py --workload $w --steps 100 --warmup 5 --no-e2e --no-cpu 2> gpurun_out/err1.txt | tee gpurun_out/bench_${w}_n1.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   value %.0f GB/s  ms/step %.4f launches %d'%(d['value'],d['ms_per_step'],d['gpu_launches']), {k:round(v['ms']*1e3,1) for k,v in d['roofline']['per_kernel'].items()})"; tail -2 gpurun_out/err1.txt
  echo "== $w 2 GPUs"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $w --steps 100 --warmup 5 --no-e2e --no-cpu 2> gpurun_out/err2.txt | tee gpurun_out/bench_${w}_n2.json | python -c "
import sys,json
for l in sys.stdin:
    if not l.startswith('{'): continue
    d=json.loads(l); print('   value %.0f GB/s  ms/step %.4f launches %d scaling %s'%(d['value'],d['ms_per_step'],d['gpu_launches'],d['scaling']), {k:round(v['ms']*1e3,1) for k,v in d['roofline']['per_kernel'].items()})"; tail -3 gpurun_out/err2.txt
done
echo "== default bench 2 GPUs (full contract)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 2> gpurun_out/err3.txt | tee gpurun_out/bench_default_n2.json | cut -c1-600; tail -3 gpurun_out/err3.txt
