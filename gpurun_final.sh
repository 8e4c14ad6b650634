python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/r2_bench1_final.json 2> gpurun_out/r2_bench1_final.err; tail -c 2600 gpurun_out/r2_bench1_final.json | head -c 1400
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench1_ref.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench1_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_bench_under_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_stream --csv --log-file gpurun_out/r2_kstream_all.csv python tests/gpu_perf.py 2048 512 1 > /dev/null 2>&1
TSB_DEBUG_PHASES=1 python tests/gpu_perf.py 2048 512 3 2> gpurun_out/r2_timeline_2048.log | tail -1
python tests/gpu_perf.py 4096 1024 3 | tail -1
python tests/gpu_perf.py 8192 1024 3 | tail -1
python tests/gpu_configs.py 2>&1 | tail -4
