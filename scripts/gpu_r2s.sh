#!/bin/bash
# Round 2: the whole GPU suite with the BreakDancer mask on the device (tests/test_gpu_bd.py; the BAM drop-in scenario now
# binds BDData::loadRegion), then the bench line and the reference arm as the driver runs them.
tag=${1:-r2s}
mkdir -p gpurun_out
O=gpurun_out/$tag
( time timeout 600 python -m pytest tests/test_gpu_bd.py -x -q ) > ${O}_pytest_bd.log 2>&1
tail -4 ${O}_pytest_bd.log
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > ${O}_pytest.log 2>&1
echo "pytest exit $?" >> ${O}_pytest.log; tail -4 ${O}_pytest.log
( time timeout 900 python bench.py --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err ) 2> ${O}_bench.time
tail -3 ${O}_bench.time
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_ref.json 2> ${O}_bench_ref.err ) 2> ${O}_bench_ref.time
tail -3 ${O}_bench_ref.time; cat ${O}_bench_ref.json | cut -c1-600
python - <<PY
import json
d=json.load(open('${O}_bench.json'))
print('main value %.2f M/s e2e %.2f parity %s cpu %s frac %.4f'%(d['value']/1e6,d['e2e']['value']/1e6,d['parity'],d.get('cpu_baseline',{}).get('value'),d['roofline']['frac']))
for s in d['secondary']:
    if 'error' in s: print('ERR',s); continue
    print(s['config']['workload'][:60],'| value %.2f M/s e2e %.2f frac %.4f parity %s cpu %s'%(s['value']/1e6,s['e2e']['value']/1e6,s['roofline']['frac'],s['parity'],s.get('cpu_baseline',{}).get('value')))
PY
