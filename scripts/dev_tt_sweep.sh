#!/bin/bash
# Development: the live single-sequence path with the persistent tracker at 512 / 384 / 256 threads per CTA (what co-resides with it).
mkdir -p gpurun_out
for tt in 512 384 256; do
  HRBF_BENCH_TRACKER_THREADS=$tt timeout -s KILL 600 python bench.py --sequences 1 --extras 0 --steps 100 > gpurun_out/tt_$tt.json 2>gpurun_out/tt_$tt.err
  echo "TT=$tt: $(python scripts/show_bench.py gpurun_out/tt_$tt.json | grep '^value')"
done
