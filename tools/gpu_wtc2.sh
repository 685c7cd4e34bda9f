#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q --tb=short -k "conv_layer and bf16" > gpurun_out/wtc2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/wtc2_tests.log
tail -12 gpurun_out/wtc2_tests.log | cut -c1-200
if grep -q "tests rc=0" gpurun_out/wtc2_tests.log; then
  L="dU4_b64 p_k3"
  VCG_NO_WSWAP=1 timeout 300 python tools/bench_conv.py $L > gpurun_out/wtc2_off.jsonl 2>&1
  timeout 300 python tools/bench_conv.py $L > gpurun_out/wtc2_on.jsonl 2>&1
  paste -d'\n' gpurun_out/wtc2_off.jsonl gpurun_out/wtc2_on.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['layer'], d['wgrad_ms'], d['wgrad_tflops'])"
fi
