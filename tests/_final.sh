python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py > gpurun_out/r2_bench1_final.json 2> gpurun_out/r2_bench1_final.err; tail -c 3000 gpurun_out/r2_bench1_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench1_ref.json 2>/dev/null; cat gpurun_out/r2_bench1_ref.json | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_bench_under_ncu.log 2>&1
python tests/gpu_progress_count.py 256 2048 2>&1 | tail -2
