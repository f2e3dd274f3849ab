#!/bin/bash
# Stages the reference's own whole-program scenarios under build/ (git-ignored, but it travels to the GPU box with the
# snapshot): the inputs live in /root/reference, which exists only in the build container.
#   build/demo : BASELINE.json configs[0] (demo/RUNME: 3 BAMs, simulated_reference.fa) + the files the pure reference binary
#                (oracle/_ref/pindel_ref) writes for it here
#   build/colo : devtools/run_regressiontests.sh:5 (COLO-829 chr20 slice, Pindel-text input) + devtools/gold_standard
set -e
cd "$(dirname "$0")/.."
R=/root/reference
mkdir -p build/demo build/colo
cp $R/demo/simulated_config.txt $R/demo/simulated_reference.fa $R/demo/simulated_reference.fa.fai $R/demo/simulated_sample_?.bam $R/demo/simulated_sample_?.bam.bai build/demo/
chmod u+w build/demo/*
( cd build/demo && ../../oracle/_ref/pindel_ref -i simulated_config.txt -f simulated_reference.fa -o exp_demo -c ALL -T 1 > exp_demo.log 2>&1
  for f in D SI INV TD LI BP RP INT_final; do gzip -f exp_demo_$f; done; rm -f exp_demo_CloseEndMapped exp_demo.log )
gzip -c $R/test/SmallTest/hs_ref_chr20.fa > build/colo/ref.fa.gz
cp $R/test/SmallTest/COLO-829_20-p_ok-1.txt.gz build/colo/reads.txt.gz
chmod u+w build/colo/*
for f in D SI INV TD LI BP CloseEndMapped; do gzip -c $R/devtools/gold_standard/colo_test.out_$f > build/colo/gold_$f.gz; done
du -sh build/demo build/colo
