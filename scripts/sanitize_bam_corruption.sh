#!/bin/bash
# scripts/bam_corruption_campaign.py with the BAM-parsing device programs' host builds under AddressSanitizer +
# UndefinedBehaviorSanitizer: damaged records behind an intact block_size chain must never make the index, the pair discovery
# or the split-read filter read outside the stream, the record table or a record.  CPU only.
#    bash scripts/sanitize_bam_corruption.sh [--rounds N] [--seed S]
set -e
cd "$(dirname "$0")/.."
E=tests/emul
for f in bam_emul sr_emul; do
   /usr/bin/g++ -O1 -g -fPIC -shared -fsanitize=address,undefined -fno-omit-frame-pointer -Wno-unknown-pragmas -x c++ -o $E/lib$f.so $E/$f.cpp
done
export ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
set +e
LD_PRELOAD="$(g++ -print-file-name=libasan.so) $(g++ -print-file-name=libubsan.so)" python scripts/bam_corruption_campaign.py "$@"
rc=$?
rm -f $E/libbam_emul.so $E/libsr_emul.so          # the next test run rebuilds the plain libraries
exit $rc
