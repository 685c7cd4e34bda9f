#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -x -k "conv_layer" > gpurun_out/k_conv.log 2>&1; echo "kernels rc=$?"
tail -12 gpurun_out/k_conv.log | cut -c1-200
timeout 300 python tools/bench_conv.py e0_b64 d5_b64 dU4_b64 dU3_b64 eD1_b64 eR_b64 2>&1 | tail -6
echo "--- VCG_NO_EPI2=1"
VCG_NO_EPI2=1 timeout 300 python tools/bench_conv.py e0_b64 dU4_b64 dU3_b64 eD1_b64 2>&1 | tail -4
