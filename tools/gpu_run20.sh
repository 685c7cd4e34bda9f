#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -x -k "conv_layer" > gpurun_out/k_conv.log 2>&1; echo "kernels rc=$?"
tail -5 gpurun_out/k_conv.log
timeout 300 python tools/bench_conv.py d5_b64 dU4_b64 e0_b64 > gpurun_out/conv_thin.log 2>&1; echo "rc=$?"; cat gpurun_out/conv_thin.log
