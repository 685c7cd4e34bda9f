#!/bin/bash
# CTA-pair convolution kernel: correctness under a timeout, then A/B timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q --tb=short -k "conv_layer and bf16" > gpurun_out/tc2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/tc2_tests.log
tail -12 gpurun_out/tc2_tests.log | cut -c1-200
if grep -q "tests rc=0" gpurun_out/tc2_tests.log; then
  L="eD2 dU2_b64"
  VCG_NO_RING=1 timeout 300 python tools/bench_conv.py $L > gpurun_out/tc2_off.jsonl 2>&1
  timeout 300 python tools/bench_conv.py $L > gpurun_out/tc2_on.jsonl 2>&1
  paste -d'\n' gpurun_out/tc2_off.jsonl gpurun_out/tc2_on.jsonl | cut -c1-140
fi
