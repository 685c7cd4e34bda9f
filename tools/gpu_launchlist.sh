#!/bin/bash
# ncu launch list (one pass, durations only) of ONE eager batch-64 training step: the second step of the run
mkdir -p gpurun_out
timeout 130 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 740 -c 760 --csv --log-file gpurun_out/r01_launches_b64_v3.csv python bench.py --global-batch 64 --steps 1 --warmup 1 --graph 0 --no-cpu-baseline > gpurun_out/ncu_launchlist.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r01_launches_b64_v3.csv
