mkdir -p gpurun_out/r02
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_backward_lanes -s 60 -c 1 -f -o gpurun_out/r02/bw_lanes_1024 python tests/dev/one_solve.py C1:1024:f64 1 6=2,9=0 > gpurun_out/r02/ncu_lanes.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_backward_staged -s 60 -c 1 -f -o gpurun_out/r02/bw_staged_1024 python tests/dev/one_solve.py C1:1024:f64 1 6=1,9=0 >> gpurun_out/r02/ncu_lanes.log 2>&1
tail -3 gpurun_out/r02/ncu_lanes.log
