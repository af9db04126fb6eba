#!/bin/bash
# light ncu pass (instruction counts + time + dram) for given workloads: args "workload:skip:count[:ENV=V,...]"
set -u
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
for spec in "$@"; do
  IFS=: read -r w skip cnt envs <<< "$spec"
  echo "== ncu-lite $w ${envs}"
  env ${envs//,/ } timeout 600 ncu --metrics $M --clock-control none -k regex:'ring_kernel|finalize' -s $skip -c $cnt --csv --log-file gpurun_out/lite_${w}.csv python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_lite_${w}.log 2>&1; echo "rc=$?"
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/lite_${w}.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[iid], r[ik][:70]),{})[r[im]]=r[iv]
for (i,k),m in d.items():
    print(i,k,' '.join('%s=%s'%(a.split('.')[0][-18:],b) for a,b in m.items()))
PY
done
