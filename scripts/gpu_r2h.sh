#!/bin/bash
# Round 2: one warp per CTA (static shared working set), vectorised replay front end, single-pass read packing.
tag=${1:-r2h}
mkdir -p gpurun_out
O=gpurun_out/$tag
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py tests/test_gpu_errors.py -x -q ) > ${O}_pytest.log 2>&1
echo "pytest exit $?" >> ${O}_pytest.log
tail -3 ${O}_pytest.log
for fm in 2048 512 100000000; do
  PG_FILTER_MIN_POS=$fm timeout 300 python bench.py --steps 5 --warmup 3 --reads 4000000 --no-cpu-baseline > ${O}_x2_f$fm.json 2> ${O}_x2_f$fm.err
  python -c "import json;d=json.load(open('${O}_x2_f$fm.json'));print('x2 filter_min',$fm,'value %.2f M/s'%(d['value']/1e6),'e2e %.2f'%(d['e2e']['value']/1e6))"
done
for fm in 2048 512; do
  PG_FILTER_MIN_POS=$fm timeout 300 python bench.py --steps 3 --warmup 3 --reads 2000000 --max-range-index 5 --no-cpu-baseline > ${O}_x5_f$fm.json 2> ${O}_x5_f$fm.err
  python -c "import json;d=json.load(open('${O}_x5_f$fm.json'));print('x5 filter_min',$fm,'value %.2f M/s'%(d['value']/1e6),'e2e %.2f'%(d['e2e']['value']/1e6))"
done
PG_FILTER_MIN_POS=2048 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pg_search_kernel -c 1 -f -o ${O}_search_x2 \
   python scripts/quick_gpu.py 500000 150 63 > ${O}_ncu_x2.log 2>&1
PG_X=5 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pg_search_kernel -c 1 -f -o ${O}_search_x5 \
   python scripts/quick_gpu.py 200000 150 63 > ${O}_ncu_x5.log 2>&1
tail -n 3 ${O}_ncu_x2.log
