#!/bin/bash
# End of round 2: ncu captures of the final search kernel (headline workload at -x 2 and -x 5) and the launch list of bench.py.
tag=${1:-r3c}
mkdir -p gpurun_out
O=gpurun_out/$tag
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pg_search_kernel -c 1 -f -o ${O}_search_x2 python scripts/quick_gpu.py 500000 150 63 > ${O}_ncu_x2.log 2>&1
PG_X=5 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pg_search_kernel -c 1 -f -o ${O}_search_x5 python scripts/quick_gpu.py 200000 150 63 > ${O}_ncu_x5.log 2>&1
tail -2 ${O}_ncu_x5.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --secondary none > ${O}_launches.log 2>&1
ls -la gpurun_out | grep $tag
