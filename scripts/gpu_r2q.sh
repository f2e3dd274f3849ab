#!/bin/bash
# Round 2, eight GPUs: bench.py under torchrun exactly as the driver launches it, every secondary workload.
tag=${1:-r2q}
mkdir -p gpurun_out
O=gpurun_out/$tag
nvidia-smi --query-gpu=index,name --format=csv > ${O}_gpu.txt
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err ) 2> ${O}_bench.time
tail -3 ${O}_bench.time; grep -E "parity|failed|Error|error" ${O}_bench.err | tail -12
python - <<PY
import json
d=json.load(open('${O}_bench.json'))
print('main value %.2f M/s e2e %.2f ms %.2f e2e_ms %.2f parity %s'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step'],d['e2e']['ms_per_step'],d['parity']))
for s in d['secondary']:
    if 'error' in s: print('ERR',s); continue
    print(s['config']['workload'][:60],'|',s['scaling'],'value %.2f M/s e2e %.2f parity %s'%(s['value']/1e6,s['e2e']['value']/1e6,s['parity']))
PY
