#!/bin/bash
# Round-2 second profiling call: refreshed launch list (staged pipeline) and ncu --set full of the non-tracker kernels of a steady-state frame.
tag=${1:-r2b}
mkdir -p gpurun_out
echo "== launch list"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 6 --warmup 3 --sequences 1 --extras 0 > gpurun_out/${tag}_launches_bench.log 2>&1
tail -c 300 gpurun_out/${tag}_launches_bench.log
echo "== ncu --set full: every kernel of one steady-state frame except the tracker"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:'^(?!.*track_persistent).*$' -s 160 -c 26 -f -o gpurun_out/${tag}_prof_frame \
    python bench.py --steps 6 --warmup 3 --sequences 1 --extras 0 > gpurun_out/${tag}_prof_frame.log 2>&1
tail -c 200 gpurun_out/${tag}_prof_frame.log
ls -la gpurun_out/${tag}_*
