#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --graph 0 --no-cpu-baseline"
$CMD > gpurun_out/plain6.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu6.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu6.log; wc -l gpurun_out/launches_r01.csv
