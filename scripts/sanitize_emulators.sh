#!/bin/bash
# The device programs' host builds (tests/emul/*.cpp: the same __host__ __device__ functions the kernels run) under
# AddressSanitizer + UndefinedBehaviorSanitizer: an out-of-bounds read of the stream, the record table or a staged tile, or a
# shift / overflow the GPU would execute silently, stops the run here.  CPU only; complements compute-sanitizer on the box
# (profiles/r1s_sanitizer.md).     bash scripts/sanitize_emulators.sh
set -e
cd "$(dirname "$0")/.."
E=tests/emul
for f in pg_emul sv_emul rp_emul bgzf_emul bam_emul boxsort_emul sr_emul bd_emul; do
   /usr/bin/g++ -O1 -g -fPIC -shared -fsanitize=address,undefined -fno-omit-frame-pointer -Wno-unknown-pragmas -x c++ -o $E/lib$f.so $E/$f.cpp
done
export ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
LD_PRELOAD="$(g++ -print-file-name=libasan.so) $(g++ -print-file-name=libubsan.so)" python -m pytest -x -q -p no:cacheprovider \
   tests/test_emulator.py tests/test_sv_emulator.py tests/test_rp_emulator.py tests/test_rp_discovery.py tests/test_bgzf_emulator.py \
   tests/test_bam_index_emulator.py tests/test_boxsort.py tests/test_sr_emulator.py tests/test_sr_fuzz.py tests/test_bd_emulator.py \
   tests/test_emulator_filter.py tests/test_emulator_fuzz.py
rc=$?
rm -f $E/lib*_emul.so          # the next test run rebuilds the plain libraries
exit $rc
