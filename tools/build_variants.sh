#!/bin/bash
# Tuning builds of the traversal kernel: tools/build_variants.sh "name:-DRK_BATCH=64 -DRK_CTAS=4" ...
# -> rakau_b200/lib/variants/librakau_b200_<name>.so (git-ignored; travels with gpurun). Select with RK_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants rakau_b200/lib/variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr"
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  ( for f in traverse sort build; do
      if [ $f == traverse ] || echo "$flags" | grep -q "RK_SORT\|RK_BUILD\|RK_PROPS\|RK_TOPO"; then
        $NV $flags -c rakau_b200/csrc/$f.cu -o build/variants/${f}_$name.o -Xptxas -v 2> build/variants/${f}_$name.log
      else cp build/$f.o build/variants/${f}_$name.o; fi
    done
    $NV -shared -o rakau_b200/lib/variants/librakau_b200_$name.so build/variants/sort_$name.o build/variants/build_$name.o build/variants/traverse_$name.o build/leapfrog.o build/capi.o build/plummer.o -lcudart_static -lpthread -ldl -lrt
    echo "$name: $(grep -A1 'traverse_kernelIfLi0ELi0' build/variants/traverse_$name.log | grep -o 'Used [0-9]* registers' | head -1) $(grep -A2 'traverse_kernelIfLi0ELi0' build/variants/traverse_$name.log | grep -o '[0-9]* bytes spill stores' | head -1)" ) &
done
wait
