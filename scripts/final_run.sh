python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -2 gpurun_out/bench_final_n1.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_final_ref.json 2>/dev/null
bash scripts/prof_steady.sh 2>&1 | tail -45
