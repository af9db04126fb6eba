#!/bin/bash
# Round 2, eight GPUs: correctness of the push exchange at 8 ranks (logs kept), the scaling lines (C2 weak with the
# C4 strong sub-record and the NUMA-aware e2e arm, C4 strong, NCCL flavours), per-CTA timelines on every rank.
set -u
mkdir -p gpurun_out
export B200NORM_PEER_TIMEOUT_S=60
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/n${N}_topo.txt 2>&1; ls /sys/devices/system/node/ | grep node >> gpurun_out/n${N}_topo.txt; nproc >> gpurun_out/n${N}_topo.txt; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective >> gpurun_out/n${N}_topo.txt 2>&1
for mode in fused deferred; do
  echo "== peer worker $mode"; timeout 300 $TR --master-port 29541 tests/peer_worker.py $mode > gpurun_out/n${N}_peer_worker_$mode.log 2>&1; echo "rc=$?"; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/n${N}_peer_worker_$mode.log | tail -2
done
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if not l.startswith('{'): continue
    d=json.loads(l); print('   value %.0f GB/s  ms/step %.4f launches %d scaling %s validated %s'%(d['value'],d['ms_per_step'],d['gpu_launches'],d['scaling'],d.get('validated')), {k:round(v['ms']*1e3,1) for k,v in d['roofline']['per_kernel'].items()}, 'e2e', d['e2e'] and (round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), d['e2e'].get('host_numa')), 'c4', d.get('c4_strong') and {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['c4_strong'].items() if k in ('ms_per_step','one_gpu_ms','efficiency','value')})
PY
}
echo "== bench default N=$N"; timeout 900 $TR --master-port 29542 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/n${N}_bench_default.json 2> gpurun_out/n${N}_bench_default.err; echo "rc=$?"; show gpurun_out/n${N}_bench_default.json; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/n${N}_bench_default.err | tail -3
for spec in "c4 peer-fused" "c4 nccl" "c2 nccl"; do
  set -- $spec
  echo "== bench $1 $2 N=$N"; timeout 600 $TR --master-port 29543 bench.py --gpus $N --workload $1 --collective $2 --steps 100 --warmup 5 --no-e2e --no-cpu --no-c4 > gpurun_out/n${N}_bench_$1_$2.json 2> gpurun_out/n${N}_err.txt; echo "rc=$?"; show gpurun_out/n${N}_bench_$1_$2.json; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/n${N}_err.txt | tail -3
done
echo "== timelines"
timeout 300 $TR --master-port 29544 scripts/timeline_peer.py c4 2> gpurun_out/n${N}_tl_err.txt | tee gpurun_out/n${N}_timeline_c4.txt
timeout 300 $TR --master-port 29545 scripts/timeline_peer.py c2 2>> gpurun_out/n${N}_tl_err.txt | tee gpurun_out/n${N}_timeline_c2.txt
grep -v "OMP_NUM\|^\*\*\*" gpurun_out/n${N}_tl_err.txt | tail -3
