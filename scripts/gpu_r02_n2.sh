#!/bin/bash
# Round 2, two GPUs: correctness of the one-shot push exchange (fused + deferred flavours against NCCL, the ATen
# oracle and across ranks), then the bench lines: default (C2 weak + the C4 strong sub-record), C4 strong, NCCL flavour.
set -u
mkdir -p gpurun_out
export B200NORM_PEER_TIMEOUT_S=30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for mode in fused deferred; do
  echo "== peer worker $mode"; timeout 300 $TR --master-port 29541 tests/peer_worker.py $mode > gpurun_out/n2_peer_worker_$mode.log 2>&1; echo "rc=$?"; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/n2_peer_worker_$mode.log | tail -3
done
echo "== pytest peer / streams"; timeout 600 python -m pytest tests/test_peer_gpu.py tests/test_streams_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if not l.startswith('{'): continue
    d=json.loads(l); print('   value %.0f GB/s  ms/step %.4f launches %d scaling %s validated %s'%(d['value'],d['ms_per_step'],d['gpu_launches'],d['scaling'],d.get('validated')), {k:round(v['ms']*1e3,1) for k,v in d['roofline']['per_kernel'].items()}, 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'c4', d.get('c4_strong') and {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['c4_strong'].items() if k in ('ms_per_step','one_gpu_ms','efficiency','value')})
PY
}
echo "== bench default N=2"; timeout 600 $TR --master-port 29542 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/n2_bench_default.json 2> gpurun_out/n2_bench_default.err; echo "rc=$?"; show gpurun_out/n2_bench_default.json; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/n2_bench_default.err | tail -3
for spec in "c2 nccl" "c4 peer-fused" "c4 nccl"; do
  set -- $spec
  echo "== bench $1 $2 N=2"; timeout 600 $TR --master-port 29543 bench.py --gpus 2 --workload $1 --collective $2 --steps 100 --warmup 5 --no-e2e --no-cpu --no-c4 > gpurun_out/n2_bench_$1_$2.json 2> gpurun_out/n2_err.txt; echo "rc=$?"; show gpurun_out/n2_bench_$1_$2.json; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/n2_err.txt | tail -3
done
echo "== bench default N=1 (same box)"; timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/n2_bench_default_n1.json 2>/dev/null; show gpurun_out/n2_bench_default_n1.json
