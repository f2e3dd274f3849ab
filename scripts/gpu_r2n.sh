#!/bin/bash
# register-cap variants side by side; racecheck of the read-pair CTA kernel on a tiny class
tag=${1:-r2n}
mkdir -p gpurun_out
timeout 600 python scripts/kernel_ab.py build/variants/libpg_*.so 2>&1 | grep -E "^libpg|Error|error" | tee gpurun_out/${tag}_ab.txt
( time PG_RP_WARP_MAX=8 timeout 500 compute-sanitizer --tool racecheck python scripts/sanitize_rp_tiny.py ) > gpurun_out/${tag}_racecheck.txt 2>&1
tail -8 gpurun_out/${tag}_racecheck.txt
