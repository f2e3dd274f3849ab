#!/bin/bash
# two GPUs: bench.py with the end-to-end path through per-rank shared-memory result segments; filter variants side by side on GPU 0
tag=${1:-r2r}
mkdir -p gpurun_out
O=gpurun_out/$tag
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err ) 2> ${O}_bench.time
tail -3 ${O}_bench.time; grep -E "failed|Error|error|Traceback" ${O}_bench.err | tail -12
python - <<PY
import json
d=json.load(open('${O}_bench.json'))
print('main value %.2f M/s e2e %.2f ms %.2f e2e_ms %.2f parity %s seen %s'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step'],d['e2e']['ms_per_step'],d['parity']['mismatches'],d['e2e'].get('records_visible_on_rank0')))
for s in d['secondary']:
    if 'error' in s: print('ERR',s); continue
    print(s['config']['workload'][:60],'|',s['scaling'],'value %.2f M/s e2e %.2f parity %s seen %s'%(s['value']/1e6,s['e2e']['value']/1e6,s['parity']['mismatches'],s['e2e'].get('records_visible_on_rank0')))
PY
ls /dev/shm | head
CUDA_VISIBLE_DEVICES=0 timeout 600 python scripts/kernel_ab.py build/variants/libpg_w1.so build/variants/libpg_w2.so 2>&1 | grep -E "^libpg|Error|error" | tee gpurun_out/${tag}_ab.txt
