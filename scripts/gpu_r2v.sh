#!/bin/bash
# Round 2: box ordering of every reporter on the device (pg_box_sort_kind) -- kernel parity, the drop-in scenarios with the DI /
# inversion reporters bound, and the kernel time of 20 000 boxes per kind beside the oracle's loops on one host thread.
tag=${1:-r2v}
mkdir -p gpurun_out
O=gpurun_out/$tag
( time timeout 900 python -m pytest tests/test_gpu_boxsort.py tests/test_gpu_dropin.py tests/test_gpu_bd.py -q ) > ${O}_pytest.log 2>&1
echo "pytest exit $?" >> ${O}_pytest.log; tail -5 ${O}_pytest.log
timeout 600 python scripts/boxsort_kind_bench.py > ${O}_boxsort_kinds.json 2> ${O}_boxsort_kinds.err; cat ${O}_boxsort_kinds.json; tail -3 ${O}_boxsort_kinds.err
