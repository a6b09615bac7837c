#!/bin/bash
# Steady-state profile (run under gpurun): (1) launch list of 4 frames after a 36-frame warm-up (map fully confident),
# (2) one `ncu --set full` capture of each heavy kernel from the same region.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
# kernels per frame vary slightly; skip a generous number of launches (36 frames x ~38 launches) then take 160
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 170 --csv --log-file gpurun_out/launches_steady.csv \
    python bench.py --steps 4 --warmup 40 > gpurun_out/bench_ncu_steady.log 2>&1
python scripts/ncu_summary.py gpurun_out/launches_steady.csv > gpurun_out/launches_steady_summary.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"predict_hrbf|curvature_gradient|depth_filter_metric|clean_flags|fuse_associate|vertex_normal_radius|track_persistent|splat_gather|pyr_pair" \
    -s 340 -c 12 -f -o gpurun_out/prof_steady python bench.py --steps 4 --warmup 40 > gpurun_out/bench_ncu_full.log 2>&1
ls -la gpurun_out/
cat gpurun_out/launches_steady_summary.txt
