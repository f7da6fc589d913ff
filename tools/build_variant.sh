#!/bin/bash
# Builds principia_b200/libpb_variant_<name>.so: the library with
# principia_b200.cu (or PB_VARIANT_SOURCE) recompiled under extra nvcc flags
# (kernel experiments; select it with PB_LIBRARY_PATH).
# usage: [PB_VARIANT_SOURCE=pb_api_adaptive.cu] tools/build_variant.sh NAME FLAGS...
set -e
name=$1; shift
cd "$(dirname "$0")/../principia_b200/csrc"
mkdir -p ../../build/variant_$name
source=${PB_VARIANT_SOURCE:-principia_b200.cu}
object=${source%.cu}.o
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC $ARCH -std=c++17 -O3 -lineinfo --fmad=false -Xcompiler -fPIC -ccbin /usr/bin/g++ \
  --expt-relaxed-constexpr -Xptxas -v "$@" -c $source \
  -o ../../build/variant_$name/$object 2> ../../build/variant_$name/ptxas.log
others=$(ls *.o | grep -v "^$object\$")
$NVCC $ARCH -shared -cudart static -o ../libpb_variant_$name.so \
  ../../build/variant_$name/$object $others
grep -A3 "SrknFlowKernelILi256ELi2ELb0" ../../build/variant_$name/ptxas.log | grep -E "Used|spill"
