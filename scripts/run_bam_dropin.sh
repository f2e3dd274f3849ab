#!/bin/bash
# One-off whole-program checks with BAM input (inputs staged in build/demo/ by hand, they live in /root/reference):
#   (1) BASELINE.json configs[0]: demo/RUNME  "pindel -i simulated_config.txt -f simulated_reference.fa -o bamtest -c ALL"
#   (2) devtools/run_regressiontests.sh:7     "pindel -f sim1chrVs2.fa -i sim1chrVs2.conf_for_test -k -s -l -T 1"
# reference host + libpindel_pg search vs the files the pure reference binary wrote for the same inputs.
set -e
cd "$(dirname "$0")/../build/demo"
export LD_LIBRARY_PATH=$PWD/../../pindel_b200:$LD_LIBRARY_PATH
ok=1
SECONDS=0
../../oracle/_ref/pindel_gpu_demo -i simulated_config.txt -f simulated_reference.fa -o gpu_demo -c ALL -T 1 > gpu_demo.log 2> gpu_demo.err || { tail -5 gpu_demo.err; exit 1; }
echo "demo/RUNME config: wall seconds $SECONDS"; grep "served by" gpu_demo.err || true
for f in D SI INV TD LI BP RP INT_final; do
  if zcat exp_demo_$f.gz | cmp -s - gpu_demo_$f; then echo "demo _$f identical to the CPU reference build"; else echo "demo _$f DIFFERS"; ok=0; fi
done
SECONDS=0
../../oracle/_ref/pindel_gpu_demo -f sim1chrVs2.fa -i sim_conf.txt -o gpu_sim -k -s -l -T 1 > gpu_sim.log 2> gpu_sim.err || { tail -5 gpu_sim.err; exit 1; }
echo "simulated_test scenario: wall seconds $SECONDS"; grep "served by" gpu_sim.err || true
for f in D SI INV TD LI BP RP INT_final CloseEndMapped; do
  if zcat exp_sim_$f.gz | cmp -s - gpu_sim_$f; then echo "simulated_test _$f identical to the CPU reference build (= devtools/gold_standard)"; else echo "simulated_test _$f DIFFERS"; ok=0; fi
done
[ $ok = 1 ] && echo "BAM-DROPIN-OK"
