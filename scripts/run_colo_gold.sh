#!/bin/bash
# Whole-program drop-in check on the reference's own regression scenario (devtools/run_regressiontests.sh:5):
# reference host + libpindel_pg search vs devtools/gold_standard/colo_test.out_*.  Inputs must be staged in build/colo/
# (they live in /root/reference/test/SmallTest, which does not exist on the GPU box).
set -e
cd "$(dirname "$0")/.."
W=$(mktemp -d)
zcat build/colo/ref.fa.gz > $W/hs_ref_chr20.fa
printf '20\t62435964\t4\t70\t71\n' > $W/hs_ref_chr20.fa.fai
zcat build/colo/reads.txt.gz > $W/reads.txt
export LD_LIBRARY_PATH=$PWD/pindel_b200:$LD_LIBRARY_PATH
SECONDS=0; oracle/_ref/pindel_gpu_demo -f $W/hs_ref_chr20.fa -p $W/reads.txt -o $W/colo_test.out -k -s -l -T 1 > $W/log.txt 2> $W/err.txt || { tail -5 $W/err.txt; exit 1; }
echo "wall seconds: $SECONDS"; grep -E "served by" $W/err.txt || true
grep -E "seconds\.$" $W/log.txt | tail -6 || true
ok=1
for f in D SI INV TD LI BP CloseEndMapped; do
  if zcat build/colo/gold_$f.gz | cmp -s - $W/colo_test.out_$f; then echo "colo_test.out_$f identical to devtools/gold_standard"; else echo "colo_test.out_$f DIFFERS"; ok=0; fi
done
[ $ok = 1 ] && echo "COLO-GOLD-OK"
