#!/bin/bash
# First hardware run of the reader's split-read filter in one gpurun call: its GPU tests on their own
# (so a failure there cannot hide behind the rest of the suite), the stage bench, a launch list and one full ncu capture of the
# pair kernel, then the whole GPU suite.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_sr_filter_round.sh <tag>'
# Everything lands in gpurun_out/<tag>_*.
tag=${1:-r3a}
mkdir -p gpurun_out
O=gpurun_out/$tag
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > ${O}_gpu.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_zz_sr_filter.py -m gpu -q ) > ${O}_sr_pytest.log 2>&1
echo "pytest exit $?" >> ${O}_sr_pytest.log
timeout 600 python scripts/sr_filter_bench.py > ${O}_sr_bench.json 2> ${O}_sr_bench.err
if [ -z "$SKIP_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pg_ --csv --log-file ${O}_sr_launches.csv \
   python scripts/sr_filter_bench.py --pairs 400000 --ref-pairs 20000 --reps 1 > ${O}_sr_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pg_sr_ -c 4 -f -o ${O}_sr_kernels \
   python scripts/sr_filter_bench.py --pairs 400000 --ref-pairs 20000 --reps 1 > ${O}_sr_ncu.log 2>&1
fi
( time timeout 900 python -m pytest tests -m gpu -x -q ) > ${O}_pytest.log 2>&1
echo "pytest exit $?" >> ${O}_pytest.log
tail -5 ${O}_sr_pytest.log; cat ${O}_sr_bench.json; tail -3 ${O}_pytest.log
