#!/bin/bash
# round-1 evidence: launch list of one eager step under ncu (shares), then --set full captures of the dominant kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu-baseline"
$CMD > gpurun_out/plain_prof.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r01b.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_r01b.csv
for spec in "eR fwd conv_tc_kernel" "eD1 fwd conv_tc_kernel" "eR dgrad conv_tc_kernel" "eR wgrad wgrad_tc_kernel" "d5 fwd conv_tc_fold" "d5 wgrad wgrad_fold"; do
  set -- $spec
  python tools/bench_one.py $1 $2 3 > gpurun_out/one_$1_$2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$3 -c 2 -o gpurun_out/r01b_$1_$2 python tools/bench_one.py $1 $2 3 > gpurun_out/ncu_$1_$2.log 2>&1
  echo "$1 $2 rc=$?"
done
ls -la gpurun_out/r01b_*.ncu-rep
