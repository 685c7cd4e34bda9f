#!/bin/bash
mkdir -p gpurun_out
python tools/bench_one.py eR fwd 3 > gpurun_out/one_eR.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 2 -o gpurun_out/r01_conv_tc_eR_fwd python tools/bench_one.py eR fwd 3 > gpurun_out/ncu_eR.log 2>&1
python tools/bench_one.py e0 fwd 3 > gpurun_out/one_e0.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 2 -o gpurun_out/r01_conv_tc_e0_fwd python tools/bench_one.py e0 fwd 3 > gpurun_out/ncu_e0.log 2>&1
python tools/bench_one.py d5 fwd 3 > gpurun_out/one_d5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_tc_rows -c 2 -o gpurun_out/r01_conv_rows_d5_fwd python tools/bench_one.py d5 fwd 3 > gpurun_out/ncu_d5.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_eR.log
