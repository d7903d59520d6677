# Round measurements on one GPU: bench (both arms), launch list of the bench command, ncu full captures,
# the larger BASELINE configs.  Usage (GPU box): bash tests/dev/run_meas.sh
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01b_launches_bench_py.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_backward -s 3 -c 2 -f -o gpurun_out/r01b_backward_f64 python tests/profile_backward.py 262144 f64 > gpurun_out/pb.log 2>&1
ncu --set full --clock-control none --import-source on -s 450 -c 5 -f -o gpurun_out/r01b_late_round python tests/profile_solve.py 4096 f64 > gpurun_out/pl2.log 2>&1
for c in "C1 4096 f64" "C1 4096 f32" "C1 65536 f64" "C1 65536 f32" "C1 262144 f64" "C2 262144 f64" "C3 131072 f64" "C4 65536 f32" "C4 524288 f32"; do python tests/run_configs.py $c >> gpurun_out/configs.jsonl 2>> gpurun_out/configs.err; done
cat gpurun_out/configs.jsonl
python tests/dev/backward_variants.py > gpurun_out/backward_variants.txt 2>&1
