#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -x -k "conv_layer" 2>&1 | tail -3
timeout 300 python tools/bench_conv.py e0_b64 dU4_b64 dU2_b64 dU1_b64 2>&1 | tail -4 | cut -c1-110
echo "--- VCG_NO_STATS_SMEM=1"
VCG_NO_STATS_SMEM=1 timeout 300 python tools/bench_conv.py e0_b64 dU4_b64 dU2_b64 dU1_b64 2>&1 | tail -4 | cut -c1-110
