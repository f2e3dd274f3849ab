#!/bin/bash
# Development tool: builds of libpindel_pg.so that differ in one compile-time choice, for scripts/kernel_ab.py.
#   bash scripts/build_variants.sh name:flags name:flags ...     e.g.  minb24:-DPG_MINB=24  f512:"-DPG_FILTER_MIN_POS=512u"
cd "$(dirname "$0")/.."
mkdir -p build/variants
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --compiler-bindir /usr/bin/g++ -shared -Xcompiler -fPIC"
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  /usr/local/cuda/bin/nvcc $F $flags -o build/variants/libpg_$name.so pindel_b200/csrc/pindel_pg.cu &
done
wait
ls -la build/variants
