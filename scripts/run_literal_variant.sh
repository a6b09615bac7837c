#!/bin/bash
# Round-2 item: build the library with the shaders' literal float-counter window loops and run the GPU parity
# suite against the oracle in the same mode.  Under gpurun:   gpurun -- 'bash scripts/run_literal_variant.sh'
# (build here first: scripts/build_variant.sh literal -DHRBF_LITERAL_WINDOWS -- build/ travels with the snapshot)
set -u
[ -f build/libhrbf_literal.so ] || scripts/build_variant.sh literal -DHRBF_LITERAL_WINDOWS
HRBF_LITERAL=1 HRBF_B200_LIB=build/libhrbf_literal.so python -m pytest tests -x -q -m gpu 2>&1 | tail -15
HRBF_B200_LIB=build/libhrbf_literal.so python bench.py --sequences 1 --steps 100 2>/dev/null | cut -c1-400
