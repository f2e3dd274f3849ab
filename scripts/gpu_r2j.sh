#!/bin/bash
# Round 2: the split-read filter's measurement round (scripts/gpu_sr_filter_round.sh: its tests, the stage bench beside the
# reference's reader, launch list + ncu of its kernels, the whole GPU suite), then the headline bench with the filter's default
# minimum ring length.
tag=${1:-r2j}
bash scripts/gpu_sr_filter_round.sh $tag
O=gpurun_out/$tag
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --secondary x5,wg > ${O}_bench.json 2> ${O}_bench.err
python - <<PY
import json
d=json.load(open('${O}_bench.json'))
print('main value %.2f M/s e2e %.2f parity %s'%(d['value']/1e6,d['e2e']['value']/1e6,d['parity']))
for s in d['secondary']:
    if 'error' in s: print('ERR',s); continue
    print(s['config']['workload'][:60],'| value %.2f M/s e2e %.2f frac %.4f parity %s'%(s['value']/1e6,s['e2e']['value']/1e6,s['roofline']['frac'],s['parity']))
PY
