#!/bin/bash
# Round 2, visit 1 (one GPU): the whole GPU suite incl. the new full-size oracle tests, the bench lines of every
# config (each validated against the oracle first), kernel-level baselines of the shapes this round works on,
# and ncu --set full captures of the kernels named in VERDICT r01 (C3 fwd / bwd, C5 bwd, LayerNorm fwd).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/v1_smi.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/v1_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/v1_smoke.log
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > gpurun_out/v1_pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/v1_pytest.log
echo "== bench default"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/v1_bench_default.json 2> gpurun_out/v1_bench_default.err; echo "rc=$?"; cut -c1-600 gpurun_out/v1_bench_default.json
for w in c1 c3 c4 c5; do
  echo "== bench $w"; timeout 900 python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu >> gpurun_out/v1_bench_all.jsonl 2>> gpurun_out/v1_bench_all.err; echo "rc=$?"
done
python - <<'PY'
import json
for l in open('gpurun_out/v1_bench_all.jsonl'):
    if not l.startswith('{'): continue
    d=json.loads(l); print(d['config']['workload'], 'validated', d['validated'], 'value %.0f'%d['value'], {k:(round(v['ms']*1e3,1), round(v['gbs'])) for k,v in d['roofline']['per_kernel'].items()}, 'frac %.3f'%d['roofline']['frac'])
PY
echo "== kbench"
timeout 900 python scripts/kbench.py copy:16384x4096:bf16 rms_fwd:16384x4096:bf16 rms_bwd:16384x4096:bf16 \
  ln_fwd:16384x4096:bf16 ln_fwd:16384x4096:bf16:B200NORM_NO_PSMEM=1 ln_fwd:32768x12288:bf16 ln_fwd:16384x8192:bf16 ln_fwd:65536x4096:bf16 ln_fwd:2048x4096:f32 ln_fwd:16384x4096:f32 \
  add_ln_fwd:16384x4096:bf16 add_rms_fwd:16384x4096:bf16 ln_bwd:16384x4096:bf16 ln_bwd:32768x12288:bf16 rms_bwd:32768x12288:bf16 \
  rms_fwd:4096x65536:bf16 rms_bwd:4096x65536:bf16 rms_bwd_dx:4096x65536:bf16 rms_bwd:4096x65536:bf16:B200NORM_CLUSTER=2 rms_bwd:4096x65536:bf16:B200NORM_CLUSTER=8 \
  rms_bwd:4096x65536:bf16:B200NORM_NO_PIPE=1 sm_fwd:65536x8192:bf16 sm_bwd:65536x8192:bf16 2>&1 | tee gpurun_out/v1_kbench.txt
echo "== ncu"
ncu_k() {  # tag regex skip spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/v1_prof_$1 -f python scripts/kbench.py $4 > gpurun_out/v1_ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
}
ncu_k c5_bwd norm_bwd_pipe 4 rms_bwd:4096x65536:bf16
ncu_k ln_fwd_12288 layernorm_fwd 4 ln_fwd:32768x12288:bf16
ncu_k ln_fwd_4096 layernorm_fwd 4 ln_fwd:16384x4096:bf16
ncu_k c3_fwd softmax_fwd 4 sm_fwd:65536x8192:bf16
ncu_k c3_bwd softmax_bwd 4 sm_bwd:65536x8192:bf16
ls -la gpurun_out/*.ncu-rep
echo "== launch list of the default bench"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v1_launches_c2.csv python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-validate > gpurun_out/v1_launches.log 2>&1; echo "rc=$?"
