#!/bin/bash
# Round 2, visit 13 (one GPU): single-exchange softmax forward over clusters (vocabulary-size rows); row-block outer kernel numbers.
set -u
mkdir -p gpurun_out
export B200NORM_PEER_TIMEOUT_S=20
echo "== tests"; timeout 1500 python -m pytest tests/test_vocab_gpu.py tests/test_parity_gpu.py tests/test_outer_gpu.py tests/test_abi.py -m gpu -q -x -p no:cacheprovider > gpurun_out/v13_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v13_pytest.log
args=""
for s in 4096x65536:bf16 4096x131072:bf16 2048x256000:bf16 18x1048576:bf16 8192x32768:bf16 4096x98304:bf16 2048x131072:f32 4096x131072:f16; do args="$args sm_fwd:$s sm_fwd:$s:B200NORM_SM_TWO_EXCH=1"; done
timeout 900 python scripts/kbench.py $args rms_fwd:4096x131072:bf16 ln_fwd:4096x131072:bf16 sm_bwd:4096x131072:bf16 2>&1 | tee gpurun_out/v13_kbench_vocab.txt
python scripts/timeline.py sm_fwd:4096x131072:bf16 rms_fwd:4096x131072:bf16 2>&1 | tee gpurun_out/v13_timeline.txt
S="100352x256:bf16 401408x64:bf16 802816x64:bf16 200704x128:f32 1605632x64:bf16 25088x1024:bf16 50176x512:bf16 100352x256:f16 70000x2048:f32 66000x4096:bf16"
args=""
for s in $S; do for op in onorm_fwd onorm_bwd; do args="$args $op:$s $op:$s:B200NORM_NO_OUTER_ROWS=1"; done; done
timeout 900 python scripts/kbench.py $args 2>&1 | tee gpurun_out/v13_kbench_rows.txt
