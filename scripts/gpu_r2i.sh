#!/bin/bash
# Round 2: the new bench.py (secondary workloads, parity field, CPU legs in worker processes) on one GPU, the pipelined
# path that leaves batch and results resident, and the whole GPU suite.
tag=${1:-r2i}
mkdir -p gpurun_out
O=gpurun_out/$tag
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > ${O}_pytest.log 2>&1
echo "pytest exit $?" >> ${O}_pytest.log
tail -4 ${O}_pytest.log
( time timeout 900 python bench.py --steps 5 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err ) 2> ${O}_bench.time
tail -3 ${O}_bench.time; tail -15 ${O}_bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > ${O}_bench_ref.json 2> ${O}_bench_ref.err ) 2> ${O}_bench_ref.time
tail -3 ${O}_bench_ref.time
python - <<PY
import json
d=json.load(open('${O}_bench.json'))
print('main value %.2f M/s e2e %.2f parity %s cpu %s'%(d['value']/1e6,d['e2e']['value']/1e6,d['parity'],d.get('cpu_baseline',{}).get('value')))
for s in d['secondary']:
    if 'error' in s: print('ERR',s); continue
    print(s['config']['workload'][:60],'| value %.2f M/s e2e %.2f frac %.4f parity %s cpu %s'%(s['value']/1e6,s['e2e']['value']/1e6,s['roofline']['frac'],s['parity'],s.get('cpu_baseline',{}).get('value')))
r=json.load(open('${O}_bench_ref.json')); print('ref arm', r['value'], r['reads_per_step_sampled'])
PY
