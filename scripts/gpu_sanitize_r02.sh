#!/bin/bash
# compute-sanitizer over the round-2 kernels (tests/sanitize_r02.py): memcheck, then racecheck and synccheck.
set -u
mkdir -p gpurun_out
echo "== native"; timeout 600 python tests/sanitize_r02.py 2>&1 | tail -3
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 8 python tests/sanitize_r02.py > gpurun_out/sanitize_r02_$tool.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|drive ok|Race reported|Invalid|hazard|Error|error" gpurun_out/sanitize_r02_$tool.log | sort | uniq -c | sort -rn | head -12
done
