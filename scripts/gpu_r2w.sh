#!/bin/bash
# Round 2: the chunked path with its per-read checks moved behind the copies (tests), then the end-to-end number of the headline
# workload for a few chunk sizes (PG_PIPE_CHUNK), everything else of bench.py switched off.
tag=${1:-r2w}
mkdir -p gpurun_out
O=gpurun_out/$tag
( time timeout 900 python -m pytest tests/test_gpu_errors.py tests/test_gpu_parity.py -q ) > ${O}_pytest.log 2>&1
echo "pytest exit $?" >> ${O}_pytest.log; tail -5 ${O}_pytest.log
for chunk in 1048576 524288 262144 2097152; do
  PG_PIPE_CHUNK=$chunk timeout 600 python bench.py --steps 10 --warmup 3 --secondary none --no-cpu-baseline > ${O}_bench_$chunk.json 2> ${O}_bench_$chunk.err
  python - <<PY
import json
d=json.load(open('${O}_bench_$chunk.json'))
print('chunk $chunk: value %.2f M/s  e2e %.2f M/s  ms %.2f  e2e_ms %.2f  parity %s'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step'],d['e2e']['ms_per_step'],d['parity']['mismatches']))
PY
done
