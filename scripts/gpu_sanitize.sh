#!/bin/bash
# compute-sanitizer over a small selection of the parity tests (memcheck + synccheck + racecheck).
set -u
mkdir -p gpurun_out
SEL='tests/test_parity_gpu.py::test_cluster_dsmem_reduction tests/test_dynamic_gpu.py::test_dynamic_every_row_written_once tests/test_peer_gpu.py::test_peer_reduce_world_of_one'
for tool in memcheck synccheck racecheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 5 python -m pytest $SEL -q -x -p no:cacheprovider -k "bf16 or every or 1000x4096" > gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|Race reported|Invalid|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -12
done
