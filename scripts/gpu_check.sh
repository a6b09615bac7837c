#!/bin/bash
# One GPU call: new-kernel tests first (under a hard timeout: a hung kernel must not hang the box), then the whole GPU suite,
# then the prints of the free-running parity tests, then the ICP reduction probes.  Output: gpurun_out/$1.log
out=gpurun_out/${1:-gpu_check}.log
mkdir -p gpurun_out
{
echo "== tile kernel + resident tracker tests"
timeout -s KILL 600 python -m pytest tests/test_gpu_odometry.py -q -m gpu -x -k "tile or resident" --timeout 300 2>&1 | tail -25
echo "== whole GPU suite"
timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -40
echo "== free-running parity prints"
timeout -s KILL 900 python -m pytest tests/test_gpu_vs_reference_shaders.py "tests/test_gpu_fusion.py::test_process_frame_sequence_matches_oracle" "tests/test_gpu_odometry.py::test_default_config_frame_agrees_iteration_by_iteration" -q -m gpu -s --timeout 600 2>&1 | grep -E "^frame|passed|failed|^E  |default conf|sensitivity"
echo "== ICP reduction probes"
timeout -s KILL 300 python scripts/probe_icp_kernels.py
echo "== tracker stamps"
timeout -s KILL 300 python scripts/dev_track_stamps.py 2>&1 | grep -v "Traceback\|^  File\|TypeError\|Exception ignored"
timeout -s KILL 300 python scripts/dev_bench_track.py 2>&1 | grep -E "^track|^trackAsync|^prep"
echo "== bench, one sequence"
timeout -s KILL 900 python bench.py --steps 100 > gpurun_out/${1:-gpu_check}_bench.json 2>/dev/null; python scripts/show_bench.py gpurun_out/${1:-gpu_check}_bench.json
} > $out 2>&1
cat $out
