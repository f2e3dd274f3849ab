#!/bin/bash
tag=${1:-r2l}
mkdir -p gpurun_out
bash scripts/gpu_r2k.sh $tag
timeout 1200 bash scripts/gpu_dropin_breakdown.sh > gpurun_out/${tag}_dropin_breakdown.txt 2>&1
cat gpurun_out/${tag}_dropin_breakdown.txt | cut -c1-600
( time timeout 600 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_parity.py -m gpu -x -q ) 2>&1 | tail -5
