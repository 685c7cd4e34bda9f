#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q --tb=short -k "conv_layer" > gpurun_out/epi3_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/epi3_tests.log | cut -c1-200
L="e0_b64 dU4_b64 p_c64k1 p_c64k3"
echo base; VCG_NO_EPI3=1 timeout 300 python tools/bench_conv.py $L 2>&1 | cut -c1-100
echo epi3_no_tstore; VCG_NO_TSTORE=1 timeout 300 python tools/bench_conv.py $L 2>&1 | cut -c1-100
echo epi3; timeout 300 python tools/bench_conv.py $L 2>&1 | cut -c1-100
