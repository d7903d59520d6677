set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3000 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01b_launches_bench_py.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_backward -s 3 -c 2 -f -o gpurun_out/r01b_backward_f64 python tests/profile_backward.py 262144 f64 > gpurun_out/pb.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_backward_staged|k_rollout_match" -s 100 -c 2 -f -o gpurun_out/r01b_lat_chains python tests/profile_solve.py 4096 f64 > gpurun_out/pl3.log 2>&1
tail -2 gpurun_out/pb.log gpurun_out/pl3.log
