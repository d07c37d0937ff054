#!/bin/bash
# Builds a variant of the engine for kernel experiments: tools/build_variant.sh NAME -DFLAG=... ...
# -> appleseed_b200/libasgpu_NAME.so (selected at run time with ASGPU_LIB=<path>).
set -e
name=$1; shift
cd "$(dirname "$0")/../appleseed_b200/csrc"
mkdir -p build_$name
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC,-ffp-contract=off,-pthread,-Wall,-Wno-unused-function"
for f in api.cu kernels.cu wavefront.cu sort.cu refine.cu lbvh.cu ploc.cu flatten.cpp tree_builder.cpp motion_bounds.cpp; do
  o=build_$name/${f%.*}.o
  case $f in
    kernels.cu|wavefront.cu) /usr/local/cuda/bin/nvcc $FLAGS "$@" -Xptxas -v -c -o $o $f 2> build_$name/${f%.*}.log & ;;
    *) if [ -f build/${f%.*}.o ]; then cp build/${f%.*}.o $o; else /usr/local/cuda/bin/nvcc $FLAGS -c -o $o $f; fi ;;
  esac
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libasgpu_$name.so build_$name/*.o -lpthread
grep -A2 "wide_kernelILb[01]ELb0ELi16ELi5ELb0ELb0ELb0E" build_$name/kernels.log | grep -E "spill|Used" || true
echo built ../libasgpu_$name.so
