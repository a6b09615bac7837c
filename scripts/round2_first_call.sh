#!/bin/bash
# The first GPU call of round 2, everything that round 1 prepared but could not run:
#   gpurun --timeout 900 -- 'bash scripts/round2_first_call.sh'
# Before the call, here: scripts/build_variant.sh literal -DHRBF_LITERAL_WINDOWS   (build/ travels with the snapshot)
set -u
mkdir -p gpurun_out
echo "== row 5: the reference's cudafuncs.cu kernels vs the oracle"
python oracle/gen_ref5_golden.py gpurun_out/ref_cudafuncs.npz 2>&1 | grep -v "oracle agrees" | tail -20
HRBF_REF5_LIVE=1 python -m pytest tests/test_oracle_vs_reference_row5.py -q -m gpu 2>&1 | tail -5
echo "== default build: GPU suite"
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== literal-window variant: GPU suite against the literal oracle, then a single-sequence bench"
bash scripts/run_literal_variant.sh
