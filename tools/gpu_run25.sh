#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -x -k "conv_layer" > gpurun_out/k_conv.log 2>&1; echo "kernels rc=$?"
tail -8 gpurun_out/k_conv.log | cut -c1-200
timeout 300 python tools/bench_conv.py e0_b64 d5_b64 dU4_b64 p_c64k1 2>&1 | tail -4
echo "--- VCG_NO_EPI3=1"
VCG_NO_EPI3=1 timeout 300 python tools/bench_conv.py e0_b64 d5_b64 dU4_b64 2>&1 | tail -3
