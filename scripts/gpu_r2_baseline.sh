#!/bin/bash
# Round-2 opening measurement: the search kernel as round 1 left it, at Pindel's default -x 2 and at -x 5, with
# source-level ncu captures of both (read here with scripts/ncu_lines.py).
#   gpurun --timeout 1200 -- 'bash scripts/gpu_r2_baseline.sh <tag>'
tag=${1:-r2f}
mkdir -p gpurun_out
O=gpurun_out/$tag
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > ${O}_gpu.txt 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > ${O}_bench_x2.json 2> ${O}_bench_x2.err
timeout 300 python bench.py --steps 3 --warmup 3 --reads 2000000 --max-range-index 5 --no-cpu-baseline > ${O}_bench_x5.json 2> ${O}_bench_x5.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pg_search_kernel -c 1 -f -o ${O}_search_x2 \
   python scripts/quick_gpu.py 500000 150 63 > ${O}_ncu_x2.log 2>&1
PG_X=5 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pg_search_kernel -c 1 -f -o ${O}_search_x5 \
   python scripts/quick_gpu.py 100000 150 63 > ${O}_ncu_x5.log 2>&1
cat ${O}_bench_x2.json ${O}_bench_x5.json; tail -3 ${O}_ncu_x2.log ${O}_ncu_x5.log
