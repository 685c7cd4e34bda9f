#!/bin/bash
mkdir -p gpurun_out
python tools/bench_one.py d5 dgrad 3 > gpurun_out/one_d5_dgrad.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 1 -o gpurun_out/r01c_d5_dgrad python tools/bench_one.py d5 dgrad 3 > gpurun_out/ncu_d5_dgrad.log 2>&1
echo rc=$?; ls -la gpurun_out/r01c_d5_dgrad.ncu-rep
