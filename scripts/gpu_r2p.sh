#!/bin/bash
# Round 2, two GPUs: the NCCL test (shard, gather from the result slots, merge, oracle), the device group on real devices, the
# drop-in over a group, bench.py under torchrun with every secondary workload.
tag=${1:-r2p}
mkdir -p gpurun_out
O=gpurun_out/$tag
nvidia-smi --query-gpu=index,name --format=csv > ${O}_gpu.txt
( time timeout 900 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_parity.py::test_device_group_equals_single_device" "tests/test_gpu_dropin.py::test_reference_host_over_a_device_group" -m gpu -x -q ) > ${O}_pytest.log 2>&1
echo "pytest exit $?" >> ${O}_pytest.log; tail -6 ${O}_pytest.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err ) 2> ${O}_bench.time
tail -3 ${O}_bench.time; grep -E "parity|failed|Error|error" ${O}_bench.err | tail -12
python - <<PY
import json
d=json.load(open('${O}_bench.json'))
print('main value %.2f M/s e2e %.2f ms %.2f e2e_ms %.2f parity %s'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step'],d['e2e']['ms_per_step'],d['parity']))
for s in d['secondary']:
    if 'error' in s: print('ERR',s); continue
    print(s['config']['workload'][:60],'|',s['scaling'],'value %.2f M/s e2e %.2f parity %s'%(s['value']/1e6,s['e2e']['value']/1e6,s['parity']))
PY
