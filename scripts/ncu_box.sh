#!/bin/bash
# On the GPU box: one `ncu --set full` capture of a kernel of a kbench spec, reduced there to text (the reports are
# 10-20 MB each and gpurun_out/ is capped at 64 MiB): raw-page CSV (every metric of the launch) and source-page
# CSV (per-line stall samples), gzip'd. usage: ncu_box.sh <tag> <kernel regex> <skip> <kbench spec> [ENV=V ...]
set -u
tag=$1; rx=$2; skip=$3; spec=$4; shift 4
mkdir -p gpurun_out /tmp/ncu
env "$@" timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o /tmp/ncu/$tag -f python scripts/kbench.py $spec > gpurun_out/ncu_$tag.log 2>&1; rc=$?
ncu -i /tmp/ncu/$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/ncu/$tag.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/ncu_${tag}_source.csv.gz
echo "ncu $tag rc=$rc $(wc -c < gpurun_out/ncu_${tag}_raw.csv) bytes raw"
