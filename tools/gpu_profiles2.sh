#!/bin/bash
# round-1 evidence (second set): ncu --set full captures of the CTA-pair kernel and the thin-N epilogue (plain run first)
mkdir -p gpurun_out
for spec in "eR fwd conv_tc2_kernel" "eR dgrad conv_tc2_kernel" "eD1 fwd conv_tc2_kernel" "e0 fwd conv_tc_kernel" "dU4 fwd conv_tc_kernel"; do
  set -- $spec
  python tools/bench_one.py $1 $2 3 > gpurun_out/one_$1_$2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$3 -c 1 -f -o gpurun_out/r01f_$1_$2 python tools/bench_one.py $1 $2 3 > gpurun_out/ncu_$1_$2.log 2>&1
  echo "$1 $2 rc=$?"
done
ls -la gpurun_out/r01f_*.ncu-rep
