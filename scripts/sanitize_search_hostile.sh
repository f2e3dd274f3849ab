#!/bin/bash
# scripts/search_hostile_campaign.py with the lane emulator (the warp program of pg_warp.cuh) built under AddressSanitizer +
# UndefinedBehaviorSanitizer: hostile per-read fields must never take the search outside the packed reference, the batch or the
# run pool.  CPU only; slow (ucontext fibers under ASan: about 5 s per round of 300 reads).
#    bash scripts/sanitize_search_hostile.sh [--rounds N] [--seed S]
cd "$(dirname "$0")/.."
export PG_EMUL_CXXFLAGS="-fsanitize=address,undefined -fno-omit-frame-pointer"
export ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
LD_PRELOAD="$(g++ -print-file-name=libasan.so) $(g++ -print-file-name=libubsan.so)" python scripts/search_hostile_campaign.py "$@"
rc=$?
rm -f tests/emul/libpg_emul.so          # the next test run rebuilds the plain library
exit $rc
