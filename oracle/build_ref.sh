#!/bin/bash
# Builds oracle/_ref/libref_reduce.so: the REFERENCE's own CUDA reduction kernels
# (Core/src/Cuda/reduce.cu + containers/device_memory.cpp), compiled UNMODIFIED from where they lie
# under /root/reference with the reference's own nvcc flags (Core/src/CMakeLists.txt:74-75), plus
# our extern "C" wrapper oracle/ref_shim.cu.  Test infrastructure only (see oracle/orc.h).
# And oracle/_ref/libref_cudafuncs.so: the reference's map / pyramid kernels (Core/src/Cuda/cudafuncs.cu), also compiled
# unmodified; its one texture<> reference (removed in CUDA 12) is served by the force-included oracle/ref_texshim.h.
# The rest of the reference's path is GLSL + Pangolin + Eigen: not buildable here.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${REF:-/root/reference/Core/src/Cuda}
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "build_ref.sh: $REF not present (GPU box?) -- keeping prebuilt $OUT"; exit 0; }
mkdir -p "$OUT"
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="--ftz=true --prec-div=false --prec-sqrt=false -O3 -Xcompiler -fPIC -w -I$REF"
if ! { [ "$OUT/libref_reduce.so" -nt "$HERE/ref_shim.cu" ] && [ "$OUT/libref_reduce.so" -nt "$REF/reduce.cu" ]; }; then
  nvcc $ARCH $FLAGS -c "$REF/reduce.cu" -o "$OUT/reduce.o"
  nvcc $ARCH $FLAGS -x cu -c "$REF/containers/device_memory.cpp" -o "$OUT/device_memory.o"
  nvcc $ARCH $FLAGS -c "$HERE/ref_shim.cu" -o "$OUT/ref_shim.o"
  nvcc $ARCH -shared -o "$OUT/libref_reduce.so" "$OUT/reduce.o" "$OUT/device_memory.o" "$OUT/ref_shim.o" -lcudart
  rm -f "$OUT"/*.o
  echo "built $OUT/libref_reduce.so"
fi
if ! { [ "$OUT/libref_cudafuncs.so" -nt "$HERE/ref_shim_cudafuncs.cu" ] && [ "$OUT/libref_cudafuncs.so" -nt "$HERE/ref_texshim.h" ] && [ "$OUT/libref_cudafuncs.so" -nt "$REF/cudafuncs.cu" ]; }; then
  nvcc $ARCH $FLAGS -include "$HERE/ref_texshim.h" -c "$REF/cudafuncs.cu" -o "$OUT/cudafuncs.o"
  nvcc $ARCH $FLAGS -x cu -c "$REF/containers/device_memory.cpp" -o "$OUT/device_memory.o"
  nvcc $ARCH $FLAGS -c "$HERE/ref_shim_cudafuncs.cu" -o "$OUT/ref_shim_cudafuncs.o"
  nvcc $ARCH -shared -o "$OUT/libref_cudafuncs.so" "$OUT/cudafuncs.o" "$OUT/device_memory.o" "$OUT/ref_shim_cudafuncs.o" -lcudart
  rm -f "$OUT"/*.o
  echo "built $OUT/libref_cudafuncs.so"
fi
