#!/bin/bash
# Builds oracle/_ref/libref_reduce.so: the REFERENCE's own CUDA reduction kernels
# (Core/src/Cuda/reduce.cu + containers/device_memory.cpp), compiled UNMODIFIED from where they lie
# under /root/reference with the reference's own nvcc flags (Core/src/CMakeLists.txt:74-75), plus
# our extern "C" wrapper oracle/ref_shim.cu.  Test infrastructure only (see oracle/orc.h).
# The reference's other CUDA file (cudafuncs.cu) uses texture<> references, removed in CUDA 12,
# and the rest of its path is GLSL + Pangolin + Eigen: not buildable here.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${REF:-/root/reference/Core/src/Cuda}
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "build_ref.sh: $REF not present (GPU box?) -- keeping prebuilt $OUT"; exit 0; }
mkdir -p "$OUT"
if [ "$OUT/libref_reduce.so" -nt "$HERE/ref_shim.cu" ] && [ "$OUT/libref_reduce.so" -nt "$REF/reduce.cu" ]; then exit 0; fi
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="--ftz=true --prec-div=false --prec-sqrt=false -O3 -Xcompiler -fPIC -w -I$REF"
nvcc $ARCH $FLAGS -c "$REF/reduce.cu" -o "$OUT/reduce.o"
nvcc $ARCH $FLAGS -x cu -c "$REF/containers/device_memory.cpp" -o "$OUT/device_memory.o"
nvcc $ARCH $FLAGS -c "$HERE/ref_shim.cu" -o "$OUT/ref_shim.o"
nvcc $ARCH -shared -o "$OUT/libref_reduce.so" "$OUT/reduce.o" "$OUT/device_memory.o" "$OUT/ref_shim.o" -lcudart
rm -f "$OUT"/*.o
echo "built $OUT/libref_reduce.so"
