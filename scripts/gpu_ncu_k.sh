#!/bin/bash
# ncu light metrics over a kbench spec list: usage gpu_ncu_k.sh <tag> <skip> <count> <kbench specs...>
set -u
tag=$1; skip=$2; cnt=$3; shift 3
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
timeout 600 ncu --metrics $M --clock-control none -k regex:'ring_kernel|finalize' -s $skip -c $cnt --csv --log-file gpurun_out/k_${tag}.csv python scripts/kbench.py "$@" > gpurun_out/k_${tag}.log 2>&1; echo "rc=$?"
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/k_${tag}.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[iid], r[ik][:70]),{})[r[im]]=r[iv]
for (i,k),m in d.items():
    print(i,k,' '.join('%s=%s'%(a.split('.')[0][-14:],b) for a,b in m.items()))
PY
