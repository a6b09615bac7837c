#!/bin/bash
# Builds oracle/_ref/libref_host.so: the reference's own Core/src/Utils/OdometryProvider.h (pose update of the Gauss-Newton loop),
# included where it lies, compiled unmodified against the minimal Eigen stand-in oracle/eigen_mini.  Test infrastructure only.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${REF_UTILS:-/root/reference/Core/src/Utils}
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "build_ref_host.sh: $REF not present -- keeping prebuilt $OUT"; exit 0; }
mkdir -p "$OUT"
if [ "$OUT/libref_host.so" -nt "$HERE/ref_shim_host.cpp" ] && [ "$OUT/libref_host.so" -nt "$HERE/eigen_mini/Eigen/Core" ] && [ "$OUT/libref_host.so" -nt "$REF/OdometryProvider.h" ]; then exit 0; fi
g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -w -I"$HERE/eigen_mini" -I"$REF" -o "$OUT/libref_host.so" "$HERE/ref_shim_host.cpp"
echo "built $OUT/libref_host.so"
