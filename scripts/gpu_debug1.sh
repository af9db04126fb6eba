#!/bin/bash
set -u
mkdir -p gpurun_out
for args in "148 8192 bwd" "296 8192 bwd" "1184 8192 bwd" "8192 8192 bwd" "8192 4096 bwd" "8192 8192 fwd" "65536 8192 fwd"; do
  echo "== $args"; timeout 60 python scripts/debug_softmax_bwd.py $args 2>&1 | tail -3
done
echo "== stages=8 groups=2"; B200NORM_GROUPS=2 timeout 60 python scripts/debug_softmax_bwd.py 8192 8192 bwd 2>&1 | tail -3
echo "== tpr=256"; B200NORM_TPR=256 timeout 60 python scripts/debug_softmax_bwd.py 8192 8192 bwd 2>&1 | tail -3
echo "== stages=5"; B200NORM_STAGES=5 timeout 60 python scripts/debug_softmax_bwd.py 8192 8192 bwd 2>&1 | tail -3
echo "== sanitizer memcheck"; timeout 300 compute-sanitizer --tool memcheck python scripts/debug_softmax_bwd.py 1184 8192 bwd 2>&1 | tail -30
echo "== sanitizer synccheck"; timeout 300 compute-sanitizer --tool synccheck python scripts/debug_softmax_bwd.py 1184 8192 bwd 2>&1 | tail -30
