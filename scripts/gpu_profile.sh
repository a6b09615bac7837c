#!/bin/bash
# Round-2 profiling call (B200_PROFILING.md recipe): launch list of a short single-sequence bench run, ncu --set full of the ICP
# reduction launch and of the persistent tracker.  Outputs under gpurun_out/ (summaries are copied to profiles/ by hand / scripts/ncu_summaries.py)
tag=${1:-r2}
mkdir -p gpurun_out
echo "== launch list"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 6 --warmup 3 --sequences 1 --extras 0 > gpurun_out/${tag}_launches_bench.log 2>&1
tail -c 300 gpurun_out/${tag}_launches_bench.log
echo "== ncu --set full: ICP reduction launch (level 0, 640x480)"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:icp_reduce_kernel -s 210 -c 2 -f -o gpurun_out/${tag}_prof_icp \
    python scripts/probe_icp_roofline.py 640 480 > gpurun_out/${tag}_prof_icp.log 2>&1
tail -c 200 gpurun_out/${tag}_prof_icp.log
echo "== ncu --set full: persistent tracker"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:track_persistent_kernel -s 6 -c 1 -f -o gpurun_out/${tag}_prof_tracker \
    python scripts/dev_bench_track.py 640 480 4 > gpurun_out/${tag}_prof_tracker.log 2>&1
tail -c 200 gpurun_out/${tag}_prof_tracker.log
ls -la gpurun_out/${tag}_*
