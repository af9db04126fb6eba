#!/bin/bash
# Full GPU parity suite, outer-softmax experiments (two half-size tiles vs one tile; store policy), bench of the
# o1 workload, ncu full of the outer kernels.
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
W=B200NORM_OUTER_W; C=B200NORM_CLUSTER; D=B200NORM_OUTER_DEPTH; P=B200NORM_OUTER_PER_SM; S=B200NORM_OUTER_SKIP_REDUCE; ST=B200NORM_OUTER_PLAIN_ST
timeout 900 python scripts/kbench.py \
  osm_fwd:128x262144:bf16 osm_fwd:128x262144:bf16:$W=256,$D=2 osm_fwd:128x262144:bf16:$ST=1 \
  osm_bwd:128x262144:bf16 osm_bwd:128x262144:bf16:$W=256,$D=2 osm_bwd:128x262144:bf16:$ST=1 \
  osm_fwd:512x65536:bf16 osm_fwd:512x65536:bf16:$W=64,$D=2 osm_fwd:512x65536:bf16:$ST=1 \
  osm_fwd:2048x16384:bf16 osm_fwd:2048x16384:bf16:$ST=1 osm_fwd:2048x16384:bf16:$W=64,$C=4,$D=2 \
  osm_fwd:8192x18432:bf16 osm_fwd:8192x18432:bf16:$ST=1 \
  osm_bwd:8192x18432:bf16 osm_bwd:8192x18432:bf16:$ST=1 \
  osm_fwd:16384x8192:bf16 osm_fwd:16384x8192:bf16:$ST=1 2>&1 | tee gpurun_out/kbench_outer7.txt
echo "== bench o1"
timeout 600 python bench.py --workload o1 --steps 100 --warmup 5 > gpurun_out/bench_o1.json 2> gpurun_out/bench_o1.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_o1.json; tail -3 gpurun_out/bench_o1.err
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:outer_softmax -s 6 -c 2 -o gpurun_out/prof_o1 -f python bench.py --workload o1 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_o1.log 2>&1; echo "ncu rc=$?"
