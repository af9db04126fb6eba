#!/bin/bash
# Round 2, visit 4 (one GPU): whole suite after the LayerNorm-pipe default / multi-group reductions / direct epilogue,
# kernel numbers of every changed path, ncu of the reduction kernels, bench lines of every config.
set -u
mkdir -p gpurun_out
export B200NORM_PEER_TIMEOUT_S=20
echo "== tests"; timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/v4_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/v4_pytest.log
echo "== kbench"
timeout 1200 python scripts/kbench.py ln_fwd:32768x12288:bf16 ln_fwd:16384x4096:bf16 ln_fwd:16384x8192:bf16 ln_fwd:65536x4096:bf16 ln_fwd:2048x4096:f32 ln_fwd:16384x4096:f32 \
  ln_fwd:16384x2048:bf16 ln_fwd:16384x2048:bf16:B200NORM_TILE_ROWS=1 ln_fwd:32768x1024:bf16 ln_fwd:32768x1024:bf16:B200NORM_TILE_ROWS=1 ln_fwd:16384x5120:bf16 ln_fwd:16384x6144:bf16 \
  add_ln_fwd:16384x4096:bf16 add_ln_fwd:16384x4096:bf16:B200NORM_TPR=256 add_ln_fwd:16384x8192:bf16 add_ln_fwd:16384x12288:bf16 \
  ln_bwd:32768x12288:bf16 ln_bwd:4096x12288:bf16 rms_bwd:16384x4096:bf16 ln_bwd:16384x4096:bf16 rms_bwd:4096x65536:bf16 \
  red_sum:16384x4096:bf16 red_max:16384x4096:bf16 colsum:16384x4096:bf16 var_mean:16384x4096:bf16 red_sum:32768x12288:bf16 red_max:32768x12288:bf16 colsum:32768x12288:bf16 var_mean:32768x12288:bf16 \
  var_mean:32768x12288:bf16:B200NORM_TPR=256 red_sum:65536x8192:bf16 var_mean:65536x8192:bf16 red_sum:16384x4096:f32 colsum:16384x4096:f32 var_mean:16384x4096:f32 red_sum:4096x65536:bf16 var_mean:4096x65536:bf16 2>&1 | tee gpurun_out/v4_kbench.txt
echo "== ncu"
bash scripts/ncu_box.sh red_sum_12288 reduce_inner 4 red_sum:32768x12288:bf16
bash scripts/ncu_box.sh var_mean_12288 layernorm_fwd 4 var_mean:32768x12288:bf16
bash scripts/ncu_box.sh ln_fwd_4096_b layernorm_fwd 4 ln_fwd:16384x4096:bf16
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/v4_bench_default.json 2> gpurun_out/v4_bench_default.err; echo "rc=$?"
for w in c1 c3 c4 c5; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu >> gpurun_out/v4_bench_all.jsonl 2>> gpurun_out/v4_bench_all.err; echo "$w rc=$?"
done
python - <<'PY'
import json
for f in ('gpurun_out/v4_bench_default.json','gpurun_out/v4_bench_all.jsonl'):
  for l in open(f):
    if not l.startswith('{'): continue
    d=json.loads(l); print(d['config']['workload'], 'validated', d['validated'], 'value %.0f'%d['value'], {k:(round(v['ms']*1e3,1), round(v['gbs'])) for k,v in d['roofline']['per_kernel'].items()}, 'frac %.3f'%d['roofline']['frac'], 'e2e', d['e2e'] and round(d['e2e']['value'],1))
PY
du -sh gpurun_out
