#!/bin/bash
# Round 2, visit 19 (one GPU): the whole GPU suite + the default bench after the row-block / dropout work.
set -u
mkdir -p gpurun_out
export B200NORM_PEER_TIMEOUT_S=20
echo "== tests"; timeout 2400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/v19_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/v19_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/v19_bench_default.json 2>gpurun_out/v19_bench_default.err; python -c "
import json; d=json.loads(open('gpurun_out/v19_bench_default.json').read().strip().splitlines()[-1]); print(d['validated'], d['value'], {k:(round(v['ms']*1e3,1), round(v['gbs'])) for k,v in d['roofline']['per_kernel'].items()}, d['roofline']['frac'], d['e2e']['value'], d.get('cpu_baseline',{}).get('value'))"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
