#!/bin/bash
# 2-GPU sanity: peer-memory tests, the default bench and the C4 strong-scaling bench at N = 2 (as the driver launches them).
set -u
mkdir -p gpurun_out
echo "== peer tests"; timeout 600 python -m pytest tests/test_peer_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
for w in c2 c4; do
  echo "== bench $w N=2"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $w --steps 100 --warmup 5 > gpurun_out/bench_${w}_n2_final.json 2> gpurun_out/bench_${w}_n2_final.err; echo "rc=$?"
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_${w}_n2_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling','gpu_launches')}, d['config']['collective'][:40], d['e2e']['value'] if d.get('e2e') else None)
"; tail -2 gpurun_out/bench_${w}_n2_final.err
done
