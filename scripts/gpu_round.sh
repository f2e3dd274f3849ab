#!/bin/bash
# One gpurun call's worth of evidence: GPU tests, smoke, both bench arms, the SV-classification bench, and ncu captures.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag>'
# Everything lands in gpurun_out/<tag>_*.
tag=${1:-r1g}
mkdir -p gpurun_out
O=gpurun_out/$tag
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > ${O}_gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > ${O}_pytest.log 2>&1
echo "pytest exit $?" >> ${O}_pytest.log
timeout 300 python __graft_entry__.py --smoke > ${O}_smoke.log 2>&1
timeout 600 python bench.py > ${O}_bench.json 2> ${O}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > ${O}_bench_ref.json 2> ${O}_bench_ref.err
timeout 600 python scripts/sv_bench.py > ${O}_sv_bench.json 2> ${O}_sv_bench.err
if [ -z "$SKIP_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pg_ --csv --log-file ${O}_sv_launches.csv \
   python scripts/sv_bench.py --reads 1000000 --cpu-sample 0 --reps 2 > ${O}_sv_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pg_sv_kernel -c 1 -f -o ${O}_sv_kernel \
   python scripts/sv_bench.py --reads 1000000 --cpu-sample 0 --reps 1 > ${O}_sv_ncu.log 2>&1
fi
tail -3 ${O}_pytest.log; cat ${O}_smoke.log | tail -2; cat ${O}_bench.json ${O}_bench_ref.json ${O}_sv_bench.json
