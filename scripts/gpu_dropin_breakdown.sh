#!/bin/bash
# Whole-program accounting: the drop-in (reference host + libpindel_pg through oracle/ref_gpu_bridge.cpp) on the
# reference's two whole-program scenarios, beside the pure reference binary with all host threads on the same box, with the
# bridge's own breakdown of where the drop-in's wall time goes.  Inputs staged by scripts/stage_dropin_inputs.sh.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_dropin_breakdown.sh > gpurun_out/r2_dropin_breakdown.txt 2>&1'
cd "$(dirname "$0")/.."
ROOT=$PWD
export LD_LIBRARY_PATH=$ROOT/pindel_b200:$LD_LIBRARY_PATH
NP=$(nproc)
echo "host threads: $NP; $(grep -m1 'model name' /proc/cpuinfo); GPU: $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1)"
W=$(mktemp -d)
# ---- demo/RUNME (BAM input, 3 samples, -c ALL) ----
cd $ROOT/build/demo
for T in 1 $NP; do
  s=$(date +%s.%N); $ROOT/oracle/_ref/pindel_ref -i simulated_config.txt -f simulated_reference.fa -o $W/ref_demo_T$T -c ALL -T $T > $W/ref_demo_T$T.log 2>&1; e=$(date +%s.%N)
  echo "demo/RUNME   pindel_ref -T $T: wall $(python3 -c "print(round($e-$s,2))") s; phases: $(grep -E ' seconds\.$' $W/ref_demo_T$T.log | tr '\n' ';')"
done
for rep in 1 2; do
  s=$(date +%s.%N); $ROOT/oracle/_ref/pindel_gpu_demo -i simulated_config.txt -f simulated_reference.fa -o $W/gpu_demo -c ALL -T $NP > $W/gpu_demo.log 2> $W/gpu_demo.err; e=$(date +%s.%N)
  echo "demo/RUNME   drop-in (run $rep): wall $(python3 -c "print(round($e-$s,2))") s"
done
grep -E "timing:|served by" $W/gpu_demo.err
ok=1; for f in D SI INV TD LI BP RP INT_final; do zcat exp_demo_$f.gz | cmp -s - $W/gpu_demo_$f || { echo "demo _$f DIFFERS"; ok=0; }; done; [ $ok = 1 ] && echo "demo/RUNME   drop-in output identical to the reference's (8 files)"
# ---- COLO-829 chr20 regression scenario (Pindel-text input) ----
cd $ROOT
zcat build/colo/ref.fa.gz > $W/hs_ref_chr20.fa
printf '20\t62435964\t4\t70\t71\n' > $W/hs_ref_chr20.fa.fai
zcat build/colo/reads.txt.gz > $W/reads.txt
for T in 1 $NP; do
  s=$(date +%s.%N); oracle/_ref/pindel_ref -f $W/hs_ref_chr20.fa -p $W/reads.txt -o $W/ref_colo_T$T -k -s -l -T $T > $W/ref_colo_T$T.log 2>&1; e=$(date +%s.%N)
  echo "colo_test    pindel_ref -T $T: wall $(python3 -c "print(round($e-$s,2))") s; phases: $(grep -E ' seconds\.$' $W/ref_colo_T$T.log | tr '\n' ';')"
done
for rep in 1 2; do
  s=$(date +%s.%N); oracle/_ref/pindel_gpu_demo -f $W/hs_ref_chr20.fa -p $W/reads.txt -o $W/gpu_colo -k -s -l -T $NP > $W/gpu_colo.log 2> $W/gpu_colo.err; e=$(date +%s.%N)
  echo "colo_test    drop-in (run $rep): wall $(python3 -c "print(round($e-$s,2))") s; phases: $(grep -E ' seconds\.$' $W/gpu_colo.log | tr '\n' ';')"
done
grep -E "timing:|served by" $W/gpu_colo.err
ok=1; for f in D SI INV TD LI BP CloseEndMapped; do zcat build/colo/gold_$f.gz | cmp -s - $W/gpu_colo_$f || { echo "colo_test.out_$f DIFFERS"; ok=0; }; done; [ $ok = 1 ] && echo "colo_test    drop-in output identical to devtools/gold_standard (7 files)"
rm -rf $W
