#!/bin/bash
# Build tuning variants of the spectrum-resident kernel next to the library (pyqms_b200/variants/*.so, selected with
# QMS_B200_LIB) -- for A/B timing on the GPU box:  tools/dense_variants.sh "-DMS_PREFETCH=0" p0 "-DMS_RECORDS=0" r0 ...
set -e
cd "$(dirname "$0")/../pyqms_b200/csrc"
mkdir -p ../variants build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC"
while [ $# -ge 2 ]; do
  defs=$1; name=$2; shift 2
  $NVCC $FLAGS $defs -c match_spectrum.cu -o build/match_spectrum_$name.o
  $NVCC $FLAGS $defs -c index.cu -o build/index_$name.o
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libqms_b200_$name.so build/ctx.o build/trees.o build/envelope.o build/index_$name.o build/match.o build/match_spectrum_$name.o build/comm.o build/planner.o -ldl
  echo built $name
done
