#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q --tb=short -k "conv_layer and bf16" > gpurun_out/tc2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/tc2_tests.log
tail -8 gpurun_out/tc2_tests.log | cut -c1-200
if grep -q "tests rc=0" gpurun_out/tc2_tests.log; then
  L="dU3_b64"
  VCG_TC2=0 timeout 300 python tools/bench_conv.py $L 2>&1 | cut -c1-140
  timeout 300 python tools/bench_conv.py $L 2>&1 | cut -c1-140
fi
