#!/bin/bash
# Steady-state profile (run under gpurun): (1) launch list of ~5 frames after a 40-frame warm-up (map fully confident),
# (2) `ncu --set full` captures of the heavy kernels from the same region and of the ICP reduction launches of the roofline probe.
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 760 -c 100 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 6 --warmup 40 > gpurun_out/bench_ncu_final.log 2>&1
python scripts/ncu_summary.py gpurun_out/launches_final.csv > gpurun_out/launches_final_summary.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"predict_hrbf|curvature_gradient|depth_filter_metric|clean_flags|fuse_associate|vertex_normal_radius|track_persistent|splat_gather|prep_all" \
    -s 250 -c 11 -f -o gpurun_out/prof_final python bench.py --steps 6 --warmup 40 > gpurun_out/bench_ncu_full2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"icp_reduce" -s 10 -c 2 -f -o gpurun_out/prof_icp_final \
    python bench.py --steps 4 --warmup 12 > gpurun_out/bench_ncu_icp.log 2>&1
ls -la gpurun_out/*.ncu-rep
cat gpurun_out/launches_final_summary.txt
