#!/bin/bash
# Round 2, visit 7 (one GPU): whole suite on the pruned instantiation tables; the lean RMSNorm backward with the
# weight unpacked in registers (C5 and the other re-read shapes); bench lines of every workload incl. r1.
set -u
mkdir -p gpurun_out
export B200NORM_PEER_TIMEOUT_S=20
echo "== tests"; timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/v7_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/v7_pytest.log
echo "== kbench"
timeout 900 python scripts/kbench.py rms_bwd:4096x65536:bf16 rms_bwd:4096x65536:bf16:B200NORM_NO_LEAN=1 rms_bwd_dx:4096x65536:bf16 \
  rms_bwd:2048x131072:bf16 rms_bwd:2048x131072:bf16:B200NORM_NO_LEAN=1 rms_bwd:16384x8192:f32 rms_bwd:16384x8192:f32:B200NORM_NO_LEAN=1 rms_bwd:8192x32768:bf16 rms_bwd:8192x32768:bf16:B200NORM_NO_LEAN=1 \
  rms_bwd:32768x12288:bf16 rms_bwd:32768x12288:bf16:B200NORM_PIPE=1 rms_bwd:32768x12288:bf16:B200NORM_PIPE=1,B200NORM_NO_LEAN=1 rms_bwd:16384x16384:bf16 rms_bwd:16384x16384:bf16:B200NORM_NO_LEAN=1 2>&1 | tee gpurun_out/v7_kbench.txt
bash scripts/ncu_box.sh c5_bwd_lean2 rms_bwd_pipe_lean 4 rms_bwd:4096x65536:bf16
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/v7_bench_default.json 2> gpurun_out/v7_bench_default.err; echo "rc=$?"
for w in c1 c3 c4 c5 r1; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu >> gpurun_out/v7_bench_all.jsonl 2>> gpurun_out/v7_bench_all.err; echo "$w rc=$?"
done
python - <<'PY'
import json
for f in ('gpurun_out/v7_bench_default.json','gpurun_out/v7_bench_all.jsonl'):
  for l in open(f):
    if not l.startswith('{'): continue
    d=json.loads(l); print(d['config']['workload'][:50], 'validated', d['validated'], 'value %.0f'%d['value'], {k:(round(v['ms']*1e3,1), round(v['gbs'])) for k,v in d['roofline']['per_kernel'].items()}, 'frac %.3f'%d['roofline']['frac'], 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
PY
tail -3 gpurun_out/v7_bench_all.err
