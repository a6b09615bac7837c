#!/bin/bash
# Steady-state profile (run under gpurun).  bench.py runs, in this order: the single-sequence pipeline (device inputs), the roofline
# probes, the single-sequence pipeline (host inputs), then the 3-sequences-per-GPU pipelines (device inputs, host inputs).
# (1) launch list of ~5 steady frames of the single-sequence run + the ICP-reduction launches of the roofline probe,
# (2) launch list of a window that reaches into the 3-sequence run (summarised from its first 256-thread tracker launch on), (3) `ncu --set full` of the persistent tracker in both shapes.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 40"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 760 -c 100 --csv --log-file gpurun_out/launches_single.csv $B > gpurun_out/bench_ncu_single.log 2>&1
python scripts/ncu_summary.py gpurun_out/launches_single.csv > gpurun_out/launches_single_summary.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 2600 --csv --log-file gpurun_out/launches_s3.csv $B > gpurun_out/bench_ncu_s3.log 2>&1
python scripts/ncu_summary.py gpurun_out/launches_s3.csv "track_persistent_kernel<256>" > gpurun_out/launches_s3_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:track_persistent -s 90 -c 5 -f -o gpurun_out/prof_tracker $B > gpurun_out/bench_ncu_tracker.log 2>&1
ls -la gpurun_out/*.ncu-rep
cat gpurun_out/launches_single_summary.txt gpurun_out/launches_s3_summary.txt
