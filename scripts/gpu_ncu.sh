#!/bin/bash
# ncu full captures: args = list of "workload:regex:skip:count[:ENV=V,...]"
set -u
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read -r w rx skip cnt envs <<< "$spec"
  echo "== ncu $w $rx"
  env ${envs//,/ } timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o gpurun_out/prof_${w}_${rx} -f python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_${w}.log 2>&1; echo "rc=$?"
done
ls -la gpurun_out/*.ncu-rep
