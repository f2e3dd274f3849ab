#!/bin/bash
# A/B: reads per step (4 M vs 10 M) and the filter's minimum ring length, headline workload only.
tag=${1:-r2k}
mkdir -p gpurun_out
O=gpurun_out/$tag
for rd in 4000000 10000000; do
 for fm in 2048 100000000; do
  PG_FILTER_MIN_POS=$fm timeout 300 python bench.py --steps 5 --warmup 3 --reads $rd --no-cpu-baseline --secondary none > ${O}_r${rd}_f$fm.json 2> ${O}_r${rd}_f$fm.err
  python -c "import json;d=json.load(open('${O}_r${rd}_f$fm.json'));print('reads',$rd,'filter_min',$fm,'value %.2f M/s'%(d['value']/1e6),'e2e %.2f'%(d['e2e']['value']/1e6), 'ms', d['ms_per_step'])"
 done
done
timeout 300 python scripts/quick_gpu.py 10000000 150 63 5 2>&1 | tail -8
