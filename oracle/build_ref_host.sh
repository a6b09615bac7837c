#!/bin/bash
# Builds oracle/_ref/libref_host.so: the reference's own Core/src/Utils/OdometryProvider.h (pose update of the Gauss-Newton loop),
# included where it lies, compiled unmodified against the minimal Eigen stand-in oracle/eigen_mini; and oracle/_ref/libref_klg.so:
# the reference's own .klg reader (GUI/src/Tools/RawLogReader.cpp + Core/src/Utils/Resolution.cpp, unmodified; Pangolin's FileExists
# and libjpeg's declarations from oracle/host_shims, zlib from the system).  Test infrastructure only.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${REF_UTILS:-/root/reference/Core/src/Utils}
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "build_ref_host.sh: $REF not present -- keeping prebuilt $OUT"; exit 0; }
mkdir -p "$OUT"
if ! { [ "$OUT/libref_host.so" -nt "$HERE/ref_shim_host.cpp" ] && [ "$OUT/libref_host.so" -nt "$HERE/eigen_mini/Eigen/Core" ] && [ "$OUT/libref_host.so" -nt "$REF/OdometryProvider.h" ]; }; then
  g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -w -I"$HERE/eigen_mini" -I"$REF" -o "$OUT/libref_host.so" "$HERE/ref_shim_host.cpp"
  echo "built $OUT/libref_host.so"
fi
CORE="$REF/.."; TOOLS="$REF/../../../GUI/src/Tools"
if [ -f "$TOOLS/RawLogReader.cpp" ] && ! { [ "$OUT/libref_klg.so" -nt "$HERE/ref_shim_klg.cpp" ] && [ "$OUT/libref_klg.so" -nt "$HERE/host_shims/jpeglib.h" ] && [ "$OUT/libref_klg.so" -nt "$TOOLS/RawLogReader.cpp" ]; }; then
  g++ -O2 -std=c++17 -fPIC -shared -w -I"$HERE/host_shims" -I"$HERE/eigen_mini" -I"$CORE" -I"$TOOLS" -o "$OUT/libref_klg.so" \
      "$HERE/ref_shim_klg.cpp" "$TOOLS/RawLogReader.cpp" "$CORE/Utils/Resolution.cpp" -lz
  echo "built $OUT/libref_klg.so"
fi
