set -x
mkdir -p gpurun_out/r02
# K5 alone, full counters (fp64 then fp32), batch 262144
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_backward -s 4 -c 2 -o gpurun_out/r02/k5_f64_262k python tests/profile_backward.py 262144 f64 > gpurun_out/r02/k5_f64.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_backward -s 4 -c 2 -o gpurun_out/r02/k5_f32_262k python tests/profile_backward.py 262144 f32 > gpurun_out/r02/k5_f32.log 2>&1
# launch list of the bench command (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02/launches_bench_py.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r02/bench_under_ncu.log 2>&1
tail -2 gpurun_out/r02/k5_f64.log gpurun_out/r02/k5_f32.log
