set -x
mkdir -p gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python tests/dev/one_solve.py C1:262144:f64 2 > gpurun_out/r02/before_c1_262k.txt 2>&1
python tests/dev/one_solve.py C2:262144:f64 2 > gpurun_out/r02/before_c2_262k.txt 2>&1
# whole-solve launch list with DRAM bytes (one solve)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02/before_launches_c1_262k.csv python tests/dev/one_solve.py C1:262144:f64 1 > gpurun_out/r02/before_launches.log 2>&1
# full counters for the step-parallel kernels of a dense early round (round 3: skip init + 2 rounds)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_cost|k_derivs|k_forward|k_sum_trials|k_decide|k_backward" -s 16 -c 7 -o gpurun_out/r02/before_fused_262k python tests/dev/one_solve.py C1:262144:f64 1 > gpurun_out/r02/before_full.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02/gputests_before.log 2>&1
tail -3 gpurun_out/r02/gputests_before.log
cat gpurun_out/r02/before_c1_262k.txt gpurun_out/r02/before_c2_262k.txt
