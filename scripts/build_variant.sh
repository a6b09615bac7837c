#!/bin/bash
# scripts/build_variant.sh NAME "-DMACRO=.." : development build of the library with other tuning macros into build/libhrbf_NAME.so
# (use with HRBF_B200_LIB=build/libhrbf_NAME.so)
set -euo pipefail
cd "$(dirname "$0")/.."
mkdir -p build/$1
for f in hrbffusion3d_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $2 -c $f -o build/$1/$(basename $f .cu).o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/libhrbf_$1.so build/$1/*.o -lcudart
echo built build/libhrbf_$1.so
