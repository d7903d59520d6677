mkdir -p gpurun_out/r02
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02/launches_bench_py_final.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02/bench_under_ncu_final.log 2>&1
python tests/summarize_launches.py gpurun_out/r02/launches_bench_py_final.csv > gpurun_out/r02/launches_bench_py_final.summary.txt 2>&1
head -20 gpurun_out/r02/launches_bench_py_final.summary.txt
