#!/bin/bash
# Round 2: mid-size batches (from 256 k reads) cut in four chunks; parity tests of the chunked path, full-size test, bench line.
tag=${1:-r2y}
mkdir -p gpurun_out
O=gpurun_out/$tag
( time timeout 900 python -m pytest tests/test_gpu_errors.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_sv.py -q ) > ${O}_pytest.log 2>&1
echo "pytest exit $?" >> ${O}_pytest.log; tail -5 ${O}_pytest.log
( time timeout 900 python bench.py --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err ) 2> ${O}_bench.time
python - <<PY
import json
d=json.load(open('${O}_bench.json'))
print('main value %.2f M/s e2e %.2f parity %s cpu %s frac %.4f'%(d['value']/1e6,d['e2e']['value']/1e6,d['parity']['mismatches'],d.get('cpu_baseline',{}).get('value'),d['roofline']['frac']))
for s in d['secondary']:
    if 'error' in s: print('ERR',s); continue
    print(s['config']['workload'][:60],'| value %.2f M/s e2e %.2f frac %.4f parity %s cpu %s'%(s['value']/1e6,s['e2e']['value']/1e6,s['roofline']['frac'],s['parity']['mismatches'],s.get('cpu_baseline',{}).get('value')))
PY
