#!/bin/bash
# kernel variants side by side; the demo drop-in with one and with all host threads
tag=${1:-r2m}
mkdir -p gpurun_out
timeout 600 python scripts/kernel_ab.py build/variants/libpg_*.so 2>&1 | grep -E "^libpg|Error|error" | tee gpurun_out/${tag}_ab.txt
cd build/demo
export LD_LIBRARY_PATH=$PWD/../../pindel_b200:$LD_LIBRARY_PATH
for T in 1 16; do
  W=$(mktemp -d)
  s=$(date +%s.%N); ../../oracle/_ref/pindel_gpu_demo -i simulated_config.txt -f simulated_reference.fa -o $W/gpu_demo -c ALL -T $T > $W/log 2> $W/err; e=$(date +%s.%N)
  echo "demo drop-in -T $T: wall $(python3 -c "print(round($e-$s,2))") s"; grep "timing:" $W/err | cut -c1-400
  for f in D SI INV TD LI BP RP INT_final; do zcat exp_demo_$f.gz | cmp -s - $W/gpu_demo_$f || echo "  -T $T: demo _$f DIFFERS"; done
  if [ $T = 16 ]; then zcat exp_demo_RP.gz | sort > $W/a; sort $W/gpu_demo_RP > $W/b; cmp -s $W/a $W/b && echo "  -T 16: _RP has the reference's lines in another order (UpdateBD's OpenMP loop appends under a critical section)"; fi
done
