#!/bin/bash
# Round 2: the whole GPU suite on the settled kernel (24 warps per SM at 80 registers), ncu captures for profiles/ (headline
# workload at -x 2 and -x 5, the 3 Gbp shape), racecheck of the read-pair CTA kernel, the full bench line.
tag=${1:-r2o}
mkdir -p gpurun_out
O=gpurun_out/$tag
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > ${O}_pytest.log 2>&1
echo "pytest exit $?" >> ${O}_pytest.log; tail -4 ${O}_pytest.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pg_search_kernel -c 1 -f -o ${O}_search_x2 python scripts/quick_gpu.py 500000 150 63 > ${O}_ncu_x2.log 2>&1
PG_X=5 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pg_search_kernel -c 1 -f -o ${O}_search_x5 python scripts/quick_gpu.py 200000 150 63 > ${O}_ncu_x5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pg_search_kernel -c 1 -f -o ${O}_search_wg python scripts/wg_probe.py 2000000 > ${O}_ncu_wg.log 2>&1
tail -2 ${O}_ncu_wg.log
( time PG_RP_WARP_MAX=8 timeout 500 compute-sanitizer --tool racecheck python scripts/sanitize_rp_tiny.py ) > ${O}_racecheck.txt 2>&1
tail -6 ${O}_racecheck.txt
( time timeout 900 python bench.py --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err ) 2> ${O}_bench.time
tail -3 ${O}_bench.time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --secondary none > ${O}_launches.log 2>&1
python - <<PY
import json
d=json.load(open('${O}_bench.json'))
print('main value %.2f M/s e2e %.2f parity %s cpu %s'%(d['value']/1e6,d['e2e']['value']/1e6,d['parity'],d.get('cpu_baseline',{}).get('value')))
for s in d['secondary']:
    if 'error' in s: print('ERR',s); continue
    print(s['config']['workload'][:60],'| value %.2f M/s e2e %.2f frac %.4f parity %s cpu %s'%(s['value']/1e6,s['e2e']['value']/1e6,s['roofline']['frac'],s['parity'],s.get('cpu_baseline',{}).get('value')))
PY
