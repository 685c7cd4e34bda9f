#!/bin/bash
# round-1 evidence: ncu --set full captures of the dominant kernels (one shape per run; plain run first)
mkdir -p gpurun_out
for spec in "eR fwd conv_tc_kernel" "eD1 fwd conv_tc_kernel" "e0 fwd conv_tc_kernel" "eR dgrad conv_tc_kernel" "eR wgrad wgrad_tc_kernel" "d5 fwd conv_tc_fold" "d5 wgrad wgrad_fold"; do
  set -- $spec
  python tools/bench_one.py $1 $2 3 > gpurun_out/one_$1_$2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$3 -c 1 -o gpurun_out/r01d_$1_$2 python tools/bench_one.py $1 $2 3 > gpurun_out/ncu_$1_$2.log 2>&1
  echo "$1 $2 rc=$?"
done
ls -la gpurun_out/r01d_*.ncu-rep
