#!/bin/bash
mkdir -p gpurun_out
python tools/bench_one.py d5 fwd 3 > gpurun_out/one_d5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_tc_fold -c 2 -o gpurun_out/r01_conv_fold_d5_fwd python tools/bench_one.py d5 fwd 3 > gpurun_out/ncu_d5.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_d5.log
