#!/bin/bash
# Round 2, end: whole GPU suite, whole-program accounting of the drop-in with every stage bound (BreakDancer mask and the DI /
# inversion reporters included), bench line and reference arm as the driver runs them.
tag=${1:-r2x}
mkdir -p gpurun_out
O=gpurun_out/$tag
( time timeout 1500 python -m pytest tests -m gpu -q ) > ${O}_pytest.log 2>&1
echo "pytest exit $?" >> ${O}_pytest.log; tail -5 ${O}_pytest.log
bash scripts/gpu_dropin_breakdown.sh > ${O}_dropin_breakdown.txt 2>&1; grep -E "wall|identical|DIFFERS" ${O}_dropin_breakdown.txt
( time timeout 900 python bench.py --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err ) 2> ${O}_bench.time
tail -3 ${O}_bench.time
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_ref.json 2> ${O}_bench_ref.err ) 2> ${O}_bench_ref.time
python - <<PY
import json
d=json.load(open('${O}_bench.json'))
print('main value %.2f M/s e2e %.2f parity %s cpu %s frac %.4f'%(d['value']/1e6,d['e2e']['value']/1e6,d['parity']['mismatches'],d.get('cpu_baseline',{}).get('value'),d['roofline']['frac']))
for s in d['secondary']:
    if 'error' in s: print('ERR',s); continue
    print(s['config']['workload'][:60],'| value %.2f M/s e2e %.2f frac %.4f parity %s cpu %s'%(s['value']/1e6,s['e2e']['value']/1e6,s['roofline']['frac'],s['parity']['mismatches'],s.get('cpu_baseline',{}).get('value')))
r=json.load(open('${O}_bench_ref.json'))
print('reference arm %.1f reads/s, %.0f ms per step'%(r['value'], r['ms_per_step']))
PY
