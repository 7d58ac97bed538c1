#!/bin/bash
# Kernel-variant experiments: build libmsim_b200 with other block sizes / occupancy targets of the staged kernel.
# usage: scripts/build_variants.sh "128 4" "256 2" "256 3" ...   -> build/variants/libmsim_<block>_<minb>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
for v in "$@"; do
  set -- $v
  out=build/variants/libmsim_$1_$2.so
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -shared \
       -DMSIM_SBLOCK=$1 -DMSIM_SMINB=$2 $EXTRA -Xptxas -v -o $out microsimulation_b200/csrc/capi.cu 2>&1 | grep -A2 "person_staged_kernel" | grep -E "registers|spill" | sed "s/^/[$1 $2] /"
done
