#!/bin/bash
# A/B experiments: tools/build_variant.sh NAME -DFLAG ...  ->  zero_b200/libzero_b200_NAME.so
# (select it at run time with ZB_LIB_PATH=zero_b200/libzero_b200_NAME.so)
set -e
name=$1; shift
out=zero_b200/_build_$name; mkdir -p $out
for f in zero_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
    -c $f -o $out/$(basename ${f%.cu}).o &
done
wait
nvcc -shared -o zero_b200/libzero_b200_$name.so $out/*.o
echo zero_b200/libzero_b200_$name.so
